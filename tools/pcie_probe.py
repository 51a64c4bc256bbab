"""Developer tool: pinned-host <-> device copy bandwidth of the box (one direction, both at once, by chunk size).
The e2e leg of bench.py is bounded by these numbers (268 MB of rays in, 201 MB of hits out per step)."""
import torch


def bw(nbytes, ms):
    return nbytes / ms / 1e6


def main():
    dev = torch.device("cuda:0")
    total = 256 << 20
    h_in = torch.empty(total, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(total, dtype=torch.uint8).pin_memory()
    d_in = torch.empty(total, dtype=torch.uint8, device=dev)
    d_out = torch.empty(total, dtype=torch.uint8, device=dev)
    s0, s1 = torch.cuda.Stream(), torch.cuda.Stream()
    for chunk in (4 << 20, 16 << 20, 64 << 20, 256 << 20):
        res = {}
        for mode in ("h2d", "d2h", "both"):
            best = 1e30
            for _ in range(4):
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                s0.wait_event(e0), s1.wait_event(e0)
                for off in range(0, total, chunk):
                    if mode in ("h2d", "both"):
                        with torch.cuda.stream(s0):
                            d_in[off:off + chunk].copy_(h_in[off:off + chunk], non_blocking=True)
                    if mode in ("d2h", "both"):
                        with torch.cuda.stream(s1):
                            h_out[off:off + chunk].copy_(d_out[off:off + chunk], non_blocking=True)
                torch.cuda.current_stream().wait_stream(s0), torch.cuda.current_stream().wait_stream(s1)
                e1.record()
                torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1))
            res[mode] = best
        print(f"chunk {chunk >> 20:4d} MiB: h2d {bw(total, res['h2d']):6.1f} GB/s  d2h {bw(total, res['d2h']):6.1f} GB/s  "
              f"both {bw(total, res['both']):6.1f} GB/s per direction ({res['both']:.2f} ms for 256 MiB each way)", flush=True)


if __name__ == "__main__":
    main()
