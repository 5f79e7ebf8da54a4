"""Runs bench.multi_device_context_check on the GPUs of the box (developer aid)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "nn-sdp_b200")]
import bench, nnsdp_b200 as nb
print(json.dumps(bench.multi_device_context_check(nb, int(sys.argv[1]) if len(sys.argv) > 1 else nb.device_count())))
