import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle as O
from soundsym_b200 import api, synth
ctx = api.Context(0)
rng = np.random.default_rng(0)
frames = int(sys.argv[1]) if len(sys.argv) > 1 else 310000
m = rng.normal(size=(frames, 12)) * np.array([20, 9, 4, 3.6, 2.3, 1.6, 1.5, 1.2, 1.1, 1, 1, 1.0])
z, _, _ = O.standardize(m[:20000])
model = O.gmm_train(z, seed=0)
def t(fn, reps=5):
    fn(); fn()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    return (time.perf_counter() - t0) / reps * 1e3
sym = ctx.symbols(m, model)
print("symbols   %.2f ms" % t(lambda: ctx.symbols(m, model)))
for depth in (3, 5):
    print("votesplit depth %d %.2f ms" % (depth, t(lambda: ctx.vote_split(sym, depth, 4))))
print("partition %.2f ms" % t(lambda: ctx.partition(m, model, 3, 4)))
l0 = ctx.launches; ctx.vote_split(sym, 3, 4); print("launches per vote_split(depth 3):", ctx.launches - l0)
