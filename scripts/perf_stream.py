"""Streaming-regime timing (GPU box): linear_cluster(w+16, window_size=w), CUDA events around the
passes; reports achieved algorithmic GB/s (2*16*2^n per measurement) and streamed GB/s."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mentpy_b200 as mb

def run(w, fuse, reps=int(os.environ.get('PERF_REPS', '3'))):
    gs = mb.templates.linear_cluster(w + 16)
    ps = mb.PatternSimulator(gs, backend="cuda-sv-stream", window_size=w, fuse=fuse)
    ang = np.random.default_rng(4).uniform(0, 2 * np.pi, w + 15)
    ps.run(ang)  # warm-up (allocations)
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        ps.run(ang)
        torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
    s = ps.simulator.last_schedule
    return best, s.algorithmic_bytes, s.streamed_bytes, len(s.passes)

if __name__ == "__main__":
    ws = [int(x) for x in (sys.argv[1].split(",") if len(sys.argv) > 1 else ["28", "30"])]
    for w in ws:
        for fuse in ([int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else (1, 2, 4, 5)):
            t, algo, streamed, npass = run(w, fuse)
            print(f"w={w} fuse={fuse}: {t*1e3:9.2f} ms/pattern  passes={npass:3d}  algorithmic {algo/t/1e9:8.1f} GB/s  streamed {streamed/t/1e9:8.1f} GB/s", flush=True)
