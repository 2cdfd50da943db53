import torch, time, numpy as np
dev = torch.device('cuda')
n = 1 << 30
pinned = torch.empty(n, dtype=torch.uint8, pin_memory=True)
pageable = torch.empty(n, dtype=torch.uint8)
view = torch.from_numpy(pinned.numpy()[: n // 2])
print('view pinned?', view.is_pinned())
d = torch.empty(n, dtype=torch.uint8, device=dev)
def t(label, fn, nbytes):
    torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print('%-30s %.1f ms  %.1f GB/s' % (label, dt * 1e3, nbytes / dt / 1e9))
for _ in range(2):
    t('H2D pinned', lambda: d.copy_(pinned, non_blocking=True), n)
    t('H2D pinned numpy view .to', lambda: view.to(dev, non_blocking=True), n // 2)
    t('H2D pageable', lambda: d.copy_(pageable), n)
    t('D2H pinned', lambda: pinned.copy_(d, non_blocking=True), n)
    t('D2H pageable (.cpu())', lambda: d[: n // 8].cpu(), n // 8)
t0 = time.perf_counter(); x = torch.empty(84 << 20, dtype=torch.uint8, pin_memory=True); print('pinned alloc 84MB ms', (time.perf_counter() - t0) * 1e3)
t0 = time.perf_counter(); x = torch.empty(1 << 30, dtype=torch.uint8, pin_memory=True); print('pinned alloc 1GB ms', (time.perf_counter() - t0) * 1e3)
