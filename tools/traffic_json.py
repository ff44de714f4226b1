"""profiles/traffic.json (DRAM bytes per time step and kernel, read by bench.py for `roofline.traffic`) from a
tools/ncu_summary.py text file:   python tools/traffic_json.py profiles/<file>.txt <time steps in the capture>"""
import json
import re
import sys

sys.path.insert(0, "tools")
from roofline_table import UNIT, parse  # noqa: E402

NAMES = {"smooth_stream_kernel": "smooth_fused", "ms_planes_kernel": "ms_segments", "events_raster_kernel": "events_raster",
         "contour_link_kernel": "contour_link", "pair_scan_kernel": "pair_scan", "streamer_touch_kernel": "streamer_touch",
         "streamer_finish_kernel": "streamer_finish", "streamer_prep_kernel": "streamer_prep", "split_events_kernel": "split_events",
         "split_raster_kernel": "split_raster"}


def main(path, steps):
    out = {"_comment": "DRAM traffic per time step (dram__bytes_read.sum + dram__bytes_write.sum of one launch over {} time "
                       "steps of the 721x1440 workload, divided by {}) from the ncu --set full capture {}; bench.py scales it "
                       "to its batch size for roofline.traffic".format(steps, steps, path)}
    for k in parse(path):
        name = re.sub(r"[<(].*", "", k["name"]).replace("void ", "").strip()
        if name not in NAMES or NAMES[name] in out or "dram__bytes_read.sum" not in k:
            continue
        rd = k["dram__bytes_read.sum"][0] * UNIT[k["dram__bytes_read.sum"][1]]
        wr = k["dram__bytes_write.sum"][0] * UNIT[k["dram__bytes_write.sum"][1]]
        out[NAMES[name]] = {"bytes_per_time_step": int((rd + wr) / steps), "source": path}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]))
