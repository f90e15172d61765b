import time, torch, os
torch.set_num_threads(os.cpu_count())
n = 8868124*81
a = torch.randint(0, 1<<24, (n,), dtype=torch.int32)
out = torch.empty(n, dtype=torch.int64).pin_memory()
for _ in range(2):
    t=time.perf_counter(); out.copy_(a); dt=time.perf_counter()-t
    print(f"widen int32->int64 {n*12/1e9:.2f} GB traffic in {dt*1e3:.1f} ms = {n*12/dt/1e9:.1f} GB/s ({os.cpu_count()} threads)")
b = torch.randint(0, 1<<24, (520753522,), dtype=torch.int64)
c = torch.empty(520753522, dtype=torch.int32).pin_memory()
for _ in range(2):
    t=time.perf_counter(); c.copy_(b); dt=time.perf_counter()-t
    print(f"narrow int64->int32 {520753522*12/1e9:.2f} GB traffic in {dt*1e3:.1f} ms = {520753522*12/dt/1e9:.1f} GB/s")
d = torch.empty_like(b).pin_memory()
for _ in range(2):
    t=time.perf_counter(); d.copy_(b); dt=time.perf_counter()-t
    print(f"memcpy int64 {520753522*16/1e9:.2f} GB traffic in {dt*1e3:.1f} ms = {520753522*16/dt/1e9:.1f} GB/s")
