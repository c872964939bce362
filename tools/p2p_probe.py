import torch, time, subprocess
print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout[:1500])
print("peer access 0->1:", torch.cuda.can_device_access_peer(0, 1))
x = torch.empty(1 << 28, dtype=torch.int64, device="cuda:0")  # 2 GB
y = torch.empty(1 << 28, dtype=torch.int64, device="cuda:1")
for _ in range(2): y.copy_(x)
torch.cuda.synchronize(0); torch.cuda.synchronize(1)
t = time.perf_counter()
for _ in range(5): y.copy_(x)
torch.cuda.synchronize(0); torch.cuda.synchronize(1)
dt = (time.perf_counter() - t) / 5
print("cudaMemcpyPeer 2 GB: %.1f GB/s" % (2.147 / dt))
