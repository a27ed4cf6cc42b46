import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import recurrent as R
from pytorch_geometric_temporal_b200.dataset import synthetic
from pytorch_geometric_temporal_b200.nn.recurrent import BatchedDCRNN
B, T = int(sys.argv[1]) if len(sys.argv) > 1 else 2, int(sys.argv[2]) if len(sys.argv) > 2 else 2
ei, ew, series = synthetic.metr_la_like(0, 64)
ei, ew, s = torch.from_numpy(ei), torch.from_numpy(ew), torch.from_numpy(series)
torch.manual_seed(0)
m = BatchedDCRNN(2, 32, 2)
X = torch.stack([s[i:i + T] for i in range(B)])
want = R.batched_dcrnn(m.state_dict(), X[:2], ei, ew)
mg = m.cuda()
with torch.no_grad():
    out = mg(X.cuda(), ei.cuda(), ew.cuda())
torch.cuda.synchronize()
print("max err", float((out[:2].cpu() - want).abs().max()), "halves", os.environ.get("STMP_DCRNN_TC_HALVES"), flush=True)
