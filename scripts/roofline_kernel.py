"""Run bench.py's roofline leg alone (the dominant kernel, timed with cold inputs).  Used as the ncu target:
  ncu --set full --clock-control none --import-source on -k regex:<kernel> -c 2 -o gpurun_out/roof python scripts/roofline_kernel.py
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench  # noqa: E402

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 64
dev = torch.device("cuda:0")
mvb, net, A, nn_ = bench.build_model(dev)
peaks = {}
p = os.path.join(ROOT, "MEASURED_PEAKS.json")
if os.path.exists(p):
    peaks = json.load(open(p))
print(json.dumps(bench.spmm_roofline(mvb, A, nn_, batch, dev, peaks, reps=10)))
