"""Diagnostic: what does a plain torch copy achieve on the streaming_llm access pattern?
(tail 508 rows + 4 sink rows of every (b,h) of a [32,32,S,80] bf16 tensor -> dense [32,32,512,80])"""
import torch, sys
dev = torch.device("cuda")
def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e-3
for S in (513, 1024, 2048, 4096):
    L = 8
    ks = [torch.randn(32, 32, S, 80, device=dev, dtype=torch.bfloat16) for _ in range(L)]
    outs = [torch.empty(32, 32, 512, 80, device=dev, dtype=torch.bfloat16) for _ in range(L)]
    def tail():
        for k, o in zip(ks, outs): o.copy_(k[:, :, -512:])
    t = timeit(tail)
    nbytes = L * 2 * 32 * 32 * 512 * 80 * 2
    print(f"S={S}: torch tail-slice copy {nbytes / t / 1e9:.0f} GB/s ({t*1e6:.0f} us for {L} layers)")
    del ks, outs
a = torch.empty(1 << 30, device=dev, dtype=torch.bfloat16); b = torch.empty_like(a)
t = timeit(lambda: b.copy_(a), 10)
print(f"plain copy 2 GiB: {2 * a.numel() * 2 / t / 1e9:.0f} GB/s")
