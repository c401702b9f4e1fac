"""Where the host time of the node path goes: FeatureTrackerNode -> PointCloud -> EstimatorNode on the rendered stream
(bench.py: node_path_closed_loop) with LFVIO_DEBUG_HOST=1, which makes the Estimator mirror print its per-stage averages."""
import os, sys
os.environ["LFVIO_DEBUG_HOST"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "lf-vio_b200"))
import bench
import lfvio
world = bench.pixel_world(int(sys.argv[1]) if len(sys.argv) > 1 else 36)
for rep in range(2):
    for store in (False, True):
        r = bench.node_path_closed_loop(lfvio, world, 0, device_store=store)
        print({k: v for k, v in r.items() if k != "path"})
