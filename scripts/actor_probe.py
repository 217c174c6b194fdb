"""GPU experiment: fused tensor-core actor vs torch (fp32 / TF32 / bf16) at BASELINE config 5's batch."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import hockey_env_b200 as hk
npz = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "td3_actors.npz")
ref = hk.load_td3_actor(npz, device="cuda:0", name="stage_3")
fused = hk.FusedActor(ref)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
obs = torch.randn((n, 18), device="cuda:0") * 2
with torch.no_grad():
    want = ref(obs)
got = fused(obs)
torch.cuda.synchronize()
print("max err", (got - want).abs().max().item(), "mean err", (got - want).abs().mean().item())
def timeit(f, reps=50):
    for _ in range(5): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
res = {"n": n}
with torch.no_grad():
    res["torch_fp32_ms"] = timeit(lambda: ref(obs))
    torch.backends.cuda.matmul.allow_tf32 = True
    res["torch_tf32_ms"] = timeit(lambda: ref(obs))
    torch.backends.cuda.matmul.allow_tf32 = False
    rb = hk.load_td3_actor(npz, device="cuda:0", name="stage_3").to(torch.bfloat16)
    ob = obs.to(torch.bfloat16)
    res["torch_bf16_ms"] = timeit(lambda: rb(ob))
res["fused_tcgen05_ms"] = timeit(lambda: fused(obs))
res["fused_tflops"] = n * 142336 / (res["fused_tcgen05_ms"] * 1e-3) / 1e12
print(json.dumps(res))
