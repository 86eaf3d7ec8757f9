"""Host<->device copy bandwidth of the box: one direction alone, and both directions at once (the ceiling of the
chunked host-buffer pipeline of tpsb_rhs_mult_host)."""
import time
import torch

n = 283_115_520  # doubles per state vector at 96^3, p = 3, 5 equations (2.26 GB)
h_in = torch.empty(n, dtype=torch.float64, pin_memory=True)
h_out = torch.empty(n, dtype=torch.float64, pin_memory=True)
d_in = torch.empty(n, dtype=torch.float64, device="cuda")
d_out = torch.empty(n, dtype=torch.float64, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
gb = n * 8 / 1e9


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


def h2d():
    with torch.cuda.stream(s1):
        d_in.copy_(h_in, non_blocking=True)


def d2h():
    with torch.cuda.stream(s2):
        h_out.copy_(d_out, non_blocking=True)


def both():
    h2d()
    d2h()


def both_chunked(chunks=32):
    m = n // chunks
    for c in range(chunks):
        with torch.cuda.stream(s1):
            d_in[c * m:(c + 1) * m].copy_(h_in[c * m:(c + 1) * m], non_blocking=True)
        with torch.cuda.stream(s2):
            h_out[c * m:(c + 1) * m].copy_(d_out[c * m:(c + 1) * m], non_blocking=True)


t = timed(h2d); print(f"H2D alone      {t * 1e3:7.1f} ms  {gb / t:6.1f} GB/s")
t = timed(d2h); print(f"D2H alone      {t * 1e3:7.1f} ms  {gb / t:6.1f} GB/s")
t = timed(both); print(f"both at once   {t * 1e3:7.1f} ms  {gb / t:6.1f} GB/s per direction")
t = timed(both_chunked); print(f"both, 32 chunks{t * 1e3:7.1f} ms  {gb / t:6.1f} GB/s per direction")
