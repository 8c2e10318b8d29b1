"""solves/s: one batch of B vs G concurrent batches of B/G (host threads, one stream each)."""
import os, sys, time, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from fast_racing_b200 import api, synth
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
probs = synth.config_problems("C4", B)
for G in [int(v) for v in (sys.argv[2:] or ["1", "2", "4"])]:
    parts = [probs[i * B // G:(i + 1) * B // G] for i in range(G)]
    batches = [api.Batch(p) for p in parts]
    res = [None] * G
    def run(i): res[i] = batches[i].optimize(1e-5)
    th = [threading.Thread(target=run, args=(i,)) for i in range(G)]
    t0 = time.perf_counter()
    for t in th: t.start()
    for t in th: t.join()
    wall = time.perf_counter() - t0
    ticks = [int(r["ticks"]) for r in res]
    ev = np.concatenate([r["evaluations"] for r in res])
    uk = res[0]["update_kernel"]
    print(f"B={B} G={G}: {B/wall:7.1f} solves/s  wall {wall:.2f}s ticks {ticks} evals mean {ev.mean():.0f} max {ev.max()}  "
          f"update kernel {1e3*uk['device_ms']/max(uk['sampled_launches'],1):.1f} us, {uk['history_bytes']/max(uk['device_ms'],1e-9)/1e6:.0f} GB/s")
    for b in batches: b.close()
