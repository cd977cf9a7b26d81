"""What the host link of this box gives a pinned copy: device->host alone, host->device alone, both at once (two streams).
The e2e leg of bench.py moves 22.9 GB out and 4.6 GB in per 10 M messages: its ceiling is the first and third line."""
import sys, torch
dev = torch.device("cuda", 0)
N = 1 << 30
d_out, d_in = torch.empty(N, dtype=torch.uint8, device=dev), torch.empty(N, dtype=torch.uint8, device=dev)
h_out, h_in = torch.empty(N, dtype=torch.uint8).pin_memory(), torch.zeros(N, dtype=torch.uint8).pin_memory()
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(d2h, h2d, reps=5, chunk=N):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    s1.wait_event(a); s2.wait_event(a)
    for _ in range(reps):
        for o in range(0, N, chunk):
            if d2h:
                with torch.cuda.stream(s1):
                    h_out[o:o + chunk].copy_(d_out[o:o + chunk], non_blocking=True)
            if h2d:
                with torch.cuda.stream(s2):
                    d_in[o:o + chunk].copy_(h_in[o:o + chunk], non_blocking=True)
    e1, e2 = torch.cuda.Event(), torch.cuda.Event()
    e1.record(s1); e2.record(s2)
    torch.cuda.current_stream().wait_event(e1); torch.cuda.current_stream().wait_event(e2)
    b.record()
    torch.cuda.synchronize()
    return reps * N / (a.elapsed_time(b) * 1e-3) / 1e9


for chunk in (N, 64 << 20, 4 << 20):
    print(f"chunk {chunk >> 20:5d} MiB  D2H alone {run(True, False, chunk=chunk):6.1f} GB/s   H2D alone {run(False, True, chunk=chunk):6.1f} GB/s   "
          f"both, per direction {run(True, True, chunk=chunk):6.1f} GB/s", flush=True)

# the same copies while the SMs stream HBM (a device-to-device copy kernel loop on a third stream): what a DMA transfer
# gets when the other slots' kernels are running
s3 = torch.cuda.Stream()
d_a, d_b = torch.empty(N, dtype=torch.uint8, device=dev), torch.empty(N, dtype=torch.uint8, device=dev)


def busy(fn, iters):
    torch.cuda.synchronize()
    with torch.cuda.stream(s3):
        for _ in range(iters):
            d_a.copy_(d_b)  # ~0.35 ms each at ~6 TB/s
    r = fn()
    torch.cuda.synchronize()
    return r


print(f"under a running HBM-bound kernel:  D2H {busy(lambda: run(True, False, reps=3), 600):6.1f} GB/s   "
      f"H2D {busy(lambda: run(False, True, reps=3), 600):6.1f} GB/s   both, per direction {busy(lambda: run(True, True, reps=3), 900):6.1f} GB/s")
