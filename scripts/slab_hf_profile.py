"""cProfile of one HF decode step with the slab cache vs a DynamicCache (host overhead only matters here)."""
import cProfile, io, os, pstats, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cs3602-llm-inference-acceleration_b200")); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import torch, kvcompress
from transformers import GPTNeoXConfig, GPTNeoXForCausalLM, DynamicCache
from kvcompress.evaluate import new_slab_for_model
torch.manual_seed(0)
cfg = GPTNeoXConfig(vocab_size=512, hidden_size=2560, num_hidden_layers=32, num_attention_heads=32, intermediate_size=2560,
                    max_position_embeddings=2048, rotary_pct=0.25)
with torch.device("cuda"):
    model = GPTNeoXForCausalLM(cfg).to(torch.bfloat16).eval()
ids = torch.randint(0, 512, (1, 600), device="cuda")
def run(mode, n=60, prof=None):
    with torch.inference_mode():
        if mode == "slab":
            slab = new_slab_for_model(model, 1, capacity=700); pkv = slab.as_hf_cache()
        else:
            pkv = DynamicCache()
        out = model(ids[:, :512], past_key_values=pkv, use_cache=True); pkv = out.past_key_values
        for t in range(512, 520): out = model(ids[:, t:t+1], past_key_values=pkv, use_cache=True)
        torch.cuda.synchronize()
        if prof: prof.enable()
        t0 = time.perf_counter()
        for t in range(520, 520 + n):
            out = model(ids[:, t:t+1], past_key_values=pkv, use_cache=True)
            if mode == "slab": slab.lengths = [520] * 32      # keep the shape fixed
            else:
                for l in pkv.layers: l.keys, l.values = l.keys[:, :, :520], l.values[:, :, :520]
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / n
        if prof: prof.disable()
    return dt
for mode in ("dynamic", "slab"):
    run(mode, 20)
    print(mode, "ms/token", round(run(mode) * 1e3, 3))
for mode in ("dynamic", "slab"):
    pr = cProfile.Profile(); run(mode, 40, pr)
    s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(12); print(mode, s.getvalue()[:2600])
