"""Pinned host <-> device copy rates on this box (the floor under bench.py's e2e): H2D, D2H, and both at once."""
import torch


def rate(fn, nbytes, iters=10):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return nbytes * iters / (e0.elapsed_time(e1) * 1e-3) / 1e9


def main():
    for mb in (5, 40, 80):
        n = mb * 1000 * 1000 // 4
        h_up, h_down = torch.empty(n).pin_memory(), torch.empty(n).pin_memory()
        d_up, d_down = torch.empty(n, device="cuda"), torch.empty(n, device="cuda")
        s2 = torch.cuda.Stream()
        up = rate(lambda: d_up.copy_(h_up, non_blocking=True), 4 * n)
        down = rate(lambda: h_down.copy_(d_down, non_blocking=True), 4 * n)

        def both():
            d_up.copy_(h_up, non_blocking=True)
            with torch.cuda.stream(s2):
                h_down.copy_(d_down, non_blocking=True)
        s2.wait_stream(torch.cuda.current_stream())
        bi = rate(both, 8 * n)
        s2.synchronize()
        print(f"{mb:3d} MB  H2D {up:6.1f} GB/s   D2H {down:6.1f} GB/s   both directions at once {bi:6.1f} GB/s (sum)", flush=True)


if __name__ == "__main__":
    main()
