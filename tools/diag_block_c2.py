import sys, numpy as np
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from oracle import pin_multi as pm
from oracle import pin_against_reference as pin
from pyblip_b200 import _lib
from pyblip_b200.linear import LinearSpikeSlab
import bench
X, y, _ = bench.make_data()
n, p = X.shape
cfg = LinearSpikeSlab(X, y)._config()
design = _lib.Design.upload(X)
S = 50
for bsize in (3, 5):
    ref, st = pm.reference_run_multi(X, y, False, pin.HP_DEFAULT, S, bsize, 0, 0, 4321)
    streams = {k: v[None] for k, v in st.items()}
    nact = (ref["betas"] != 0).sum(1)
    for mode in (_lib.MODE_GRAM, _lib.MODE_RESIDUAL):
        smp = _lib.sample_spikeslab(design, cfg, y=y, chains=1, n_keep=S, burn=0, seed=pin.TN_SEED, streams=streams, mode=mode, bsize=bsize)
        b = smp.dense(); smp.close()
        r = ref["betas"]
        err = np.abs(b - r)
        own = err / np.maximum(np.abs(r), 1e-300)
        rowmax = np.abs(r).max(1)
        print(f"bsize {bsize} mode {mode}")
        for s in list(range(0, 12)) + list(range(12, S, 6)):
            i = np.argmax(own[s])
            blk = (i // bsize) * bsize
            print(f"  sweep {s:2d} nact {nact[s]:5d} max own {own[s].max():.2e} at |ref| {abs(r[s,i]):.2e} (block max {np.abs(r[s,blk:blk+bsize]).max():.2e}) max abs err {err[s].max():.2e} rowmax {rowmax[s]:.2f}")
