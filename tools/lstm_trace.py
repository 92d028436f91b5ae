"""Print the per-step timeline of the K-split LSTM layer kernel from a TB200_LSTM_TRACE dump."""
import sys
import numpy as np
T = int(sys.argv[2]) if len(sys.argv) > 2 else 100
a = np.fromfile(sys.argv[1], dtype=np.uint64).reshape(-1, T, 8).astype(np.int64)
G = a.shape[0]
names = ["barrier seen (producer)", "full[0] seen (mma)", "full[1] seen (mma)", "tmem_full seen (epi)", "push issued (epi)",
         "recv complete (epi)", "h stored (epi)", "counter bumped (epi)"]
# reference point of step t: the LAST CTA's counter bump of step t-1
for t in (5, 50, 90):
    ref = a[:, t - 1, 7].max()
    print(f"step {t}: period {(a[:, t, 7].max() - ref)} ns")
    for s in range(8):
        v = a[:, t, s] - ref
        print(f"   {names[s]:28s} min {v.min():6d}  median {int(np.median(v)):6d}  max {v.max():6d} ns")
if len(sys.argv) > 3 and sys.argv[3] == "h16":
    print("(h16 kernel: slot 'full[1] seen' = epilogue rendezvous passed, before the gpu-scope fence)")
per = np.diff(a[:, 1:, 7].max(axis=0))
print("median step period", np.median(per), "ns")
