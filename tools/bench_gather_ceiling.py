#!/usr/bin/env python
"""What the memory system gives a plain random row gather, as a yardstick for the fused kernels on
small rows: out[i] = table[idx[i]] with idx = randperm (every row once) vs idx = arange, for rows
of 64 .. 1024 bytes.  Uses torch.index_select (library gather, fp32 columns) -- a reference point,
not part of the product path.  Prints one JSON line per row size: G rows/s and GB/s (read + write).
"""
import json
import sys
import torch

def main():
    dev = torch.device("cuda", 0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    gen = torch.Generator(device=dev).manual_seed(0)
    for row_bytes in (64, 128, 256, 512, 1024):
        C = row_bytes // 4
        V = (2 << 30) // row_bytes          # 2 GiB table
        x = torch.randn(V, C, device=dev, generator=gen)
        out = torch.empty_like(x)
        res = {"row_bytes": row_bytes, "rows": V}
        for mode in ("randperm", "arange"):
            idx = torch.randperm(V, device=dev, generator=gen) if mode == "randperm" else torch.arange(V, device=dev)
            for _ in range(2):
                torch.index_select(x, 0, idx, out=out)
            ts = []
            for _ in range(5):
                flush.zero_()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                torch.index_select(x, 0, idx, out=out)
                b.record()
                torch.cuda.synchronize()
                ts.append(a.elapsed_time(b))
            ms = sorted(ts)[len(ts) // 2]
            res[mode] = {"ms": round(ms, 4), "grows_per_s": round(V / ms / 1e6, 2),
                         "gbs_read_plus_write": round(2 * V * row_bytes / ms / 1e6, 1)}
        print(json.dumps(res), flush=True)
        del x, out

if __name__ == "__main__":
    main()
