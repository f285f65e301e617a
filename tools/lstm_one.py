"""One BiLSTM training step for ncu (kernel filter: lstm_seq): python tools/lstm_one.py [batch] [steps]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pitchextractor_b200 import JDCNet
B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2
cfg = dict(model_type="bilstm", num_layers=4, dropout=0.1, nhead=8, dim_feedforward=1536, max_len=2048)
torch.manual_seed(0)
m = JDCNet(num_class=1, sequence_model_config=cfg).cuda()
m.engine.use_graph = False
g = torch.Generator(device="cuda").manual_seed(1)
mel = torch.randn(B, 1, 80, 192, device="cuda", generator=g)
f0 = torch.rand(B, 192, device="cuda", generator=g) * 200 + 100
sil = (torch.rand(B, 192, device="cuda", generator=g) < 0.2).float()
for _ in range(n):
    out = m.engine.train_step(mel, f0, sil, 0.1)
torch.cuda.synchronize()
print("ok", out.tolist())
