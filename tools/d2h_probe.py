"""Where does the multi-GPU end-to-end figure saturate?  One process drives k GPUs at once, each
copying a 1 GiB device buffer into its own page-locked host buffer (and back), k = 1, 2, 4, 8 and a
few pairs: aggregate GB/s shows whether the ceiling is per GPU link (scales with k) or a shared
host resource (flat).  `python tools/d2h_probe.py` on a multi-GPU box."""
import json
import time

import torch

GIB = 1 << 30


def run(devs, direction, reps=4):
    dev_bufs = [torch.empty(GIB, dtype=torch.uint8, device=f"cuda:{d}") for d in devs]
    host_bufs = [torch.empty(GIB, dtype=torch.uint8, pin_memory=True) for _ in devs]
    streams = [torch.cuda.Stream(f"cuda:{d}") for d in devs]

    def go():
        for d, db, hb, st in zip(devs, dev_bufs, host_bufs, streams):
            with torch.cuda.stream(st):
                if direction == "d2h":
                    hb.copy_(db, non_blocking=True)
                else:
                    db.copy_(hb, non_blocking=True)

    go()
    for d in devs:
        torch.cuda.synchronize(d)
    t0 = time.perf_counter()
    for _ in range(reps):
        go()
    for d in devs:
        torch.cuda.synchronize(d)
    sec = time.perf_counter() - t0
    return len(devs) * reps * GIB / sec / 1e9


def main():
    n = torch.cuda.device_count()
    sets = [[0]] + [list(range(k)) for k in (2, 4, 8) if k <= n]
    if n >= 8:
        sets += [[0, 4], [0, 7], [2, 3], [4, 5, 6, 7]]
    import sys

    for direction in (sys.argv[1:] or ["d2h", "h2d"]):
        for devs in sets:
            gbs = run(devs, direction)
            print(json.dumps({"direction": direction, "gpus": devs, "aggregate_GBps": round(gbs, 1),
                              "per_gpu_GBps": round(gbs / len(devs), 1)}), flush=True)


if __name__ == "__main__":
    main()
