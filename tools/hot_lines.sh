mkdir -p /tmp/ncu
for k in events_raster streamer_touch contour_link pair_scan; do
  timeout 200 ncu --set full --import-source on --clock-control none -k regex:$k -s 1 -c 1 -o /tmp/ncu/$k python tools/ncu_kernels.py 148 > /dev/null 2>&1
  ncu -i /tmp/ncu/$k.ncu-rep --page source --csv > /tmp/ncu/$k.csv 2>/dev/null
  echo "=== $k"; python tools/ncu_hot_lines.py /tmp/ncu/$k.csv 45
done > gpurun_out/r2q_hot_lines.txt 2>&1
