import sys, json
sys.path.insert(0, "/root/repo")
import torch, bench
print(json.dumps(bench.hbm_kernels(torch.device("cuda:0")), indent=1))
