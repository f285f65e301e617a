"""One log-mel workload for ncu: python tools/logmel_one.py [batch] [seconds] [repeats]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pitchextractor_b200.mel import LogMel
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
L = int(float(sys.argv[2]) * 24000) // 4 * 4 if len(sys.argv) > 2 else 240000
n = int(sys.argv[3]) if len(sys.argv) > 3 else 3
lm = LogMel("cuda")
w = torch.randn(B, L, device="cuda") * 0.1
for _ in range(n):
    y = lm(w)
torch.cuda.synchronize()
print("ok", tuple(y.shape), float(y.mean()))
