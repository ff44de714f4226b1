"""CPU oracle for the WaveBreaking per-time-step detection path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``wavebreaking_b200/`` may import this
package; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs do, and there only as the checker /
reported CPU baseline, never as the product path.

It restates, function by function, the reference pipeline
(skaderli/WaveBreaking v0.3.8, cited as ``<file>:<line>`` relative to the
reference checkout) with numpy / scipy / scikit-learn / pandas only:

* where the reference calls ``scipy.ndimage.convolve`` and
  ``sklearn.metrics.DistanceMetric("haversine")`` the oracle calls the very
  same routines (both are installed in this image);
* ``skimage.measure.find_contours`` (unpinned dependency, not installed) is
  restated from its published algorithm in :mod:`oracle.skimage_contours`;
* the shapely / GEOS / geopandas pieces (``touches``, ``intersects``,
  ``buffer``+``sjoin(contains)``, the meridian split, ``intersection.area``)
  are restated on the integer lattice in :mod:`oracle.geom`.

PARITY PINNING: the reference cannot be imported here (xarray, geopandas,
shapely, scikit-image are absent and there is no network) and its only fixture
``tests/data/demo_data.nc`` is missing from the checkout.  The oracle is pinned
against (i) the reference's own fixture-free known-answer tests
(``tests/test_wavebreaking.py:101-104,116-128,130-144``), (ii) live outputs of
the real scipy / sklearn routines, (iii) scikit-image's published vectors and the
Shapely manual's predicate examples, and (iv) an evaluator written from the DE-9IM
definitions (``tests/test_predicates.py``).  Everything that depends on skimage /
GEOS semantics beyond those vectors is therefore **parity unpinned** and says so
in DESIGN.md; ``tools/make_reference_fixtures.py`` produces the fixture that pins
it wherever the real package can be installed (``tests/test_reference_fixtures.py``).

The meridian split (:func:`oracle.geom.split_ring_at_meridian`, a planar-graph face
walk) and the overlap decision of ``track_events``
(:func:`oracle.geom.overlap_positive_exact`, a rational slab sweep) are formulated
independently of the product's algorithms (chain clipper; crossing / touching rules).
"""

from . import skimage_contours, geom, pipeline  # noqa: F401
