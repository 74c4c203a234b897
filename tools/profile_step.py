"""One eager, single-stream train step of bench.py's workload between cudaProfilerStart/Stop, for ncu:

  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file L.csv \\
      python tools/profile_step.py                      # launch list of the whole step
  ncu --profile-from-start off --set full --clock-control none --import-source on \\
      -k regex:"conv3d_wgrad_tc_kernel|conv3d_tc_dfold_kernel|conv3d_tc_kernel" -o R python tools/profile_step.py

Numbers printed under a profiler are not bench values."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("FPL_CUDA_GRAPH", "0")
import torch

import bench


def main():
    torch.cuda.set_device(0)
    agent = bench.build_agent("train", 1)
    agent.use_cuda_graph = False
    agent.dual_stream = False
    agent.net.wgrad_side_stream = False
    host = [bench.make_batch(11, bench.BATCH, bench.PATCH, False, True, True), bench.make_batch(12, bench.BATCH, bench.PATCH, True, True, True)]
    dev = [{k: (v.to(agent.device) if torch.is_tensor(v) else v) for k, v in b.items()} for b in host]
    for _ in range(int(os.environ.get("WARM", "2"))):
        agent.train_step(dev)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    agent.train_step(dev)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    if os.environ.get("FILTER"):
        pass


if __name__ == "__main__":
    main()
