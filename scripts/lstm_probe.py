"""cuDNN nn.LSTM(64,64) forward+backward time vs batch size (seq 50): the PPO 'EMB' LSTM runs with batch = nodes."""
import torch, time
dev = torch.device("cuda")
lstm = torch.nn.LSTM(64, 64).to(dev)
for batch in (975, 3900, 7800, 15600, 31200, 62400):
    x = torch.randn(50, batch, 64, device=dev, requires_grad=True)
    h = (torch.zeros(1, batch, 64, device=dev), torch.zeros(1, batch, 64, device=dev))
    for _ in range(2):
        out, _ = lstm(x, h); out.sum().backward()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(3):
        out, _ = lstm(x, h); out.sum().backward()
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / 3 * 1e3
    print("batch %6d: %8.2f ms fwd+bwd  (%.2f us per sequence)" % (batch, ms, ms * 1e3 / batch))
