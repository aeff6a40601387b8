import torch, time
dev = torch.device("cuda:0")
n, m = 2449029, 123700000
g = torch.Generator(device=dev).manual_seed(1)
idx = torch.randint(0, n, (m,), device=dev, generator=g, dtype=torch.int64)
ones = torch.ones(m, dtype=torch.int32, device=dev)
cnt = torch.zeros(n, dtype=torch.int32, device=dev)
for rep in range(3):
    cnt.zero_(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); cnt.index_add_(0, idx, ones); b.record(); torch.cuda.synchronize()
    print("index_add_ int32 (RED.ADD) 123.7M random into 2.45M bins:", round(a.elapsed_time(b), 3), "ms")
# scattered 4-byte stores (the scatter pass): out[perm[i]] = i
perm = torch.randperm(m, device=dev)
src = torch.arange(m, dtype=torch.int32, device=dev)
out = torch.empty(m, dtype=torch.int32, device=dev)
for rep in range(3):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); out[perm] = src; b.record(); torch.cuda.synchronize()
    print("random 4-byte scatter of 123.7M:", round(a.elapsed_time(b), 3), "ms")
# row-local scatter: destination = row start + small random offset (what the bucket build does)
rows = torch.sort(idx).values
