"""Energy per kernel class (VERDICT weak #8: every run sits at sw_power_cap, so J per step -- not time -- is what a change
has to lower).  Each class runs back to back for ~2 s at the bench shapes (B = 16, ViT-H) while NVML's energy counter
(nvmlDeviceGetTotalEnergyConsumption, mJ) and the SM clock are read around / during the loop:
    python tools/gpu_energy.py [fp16|bf16] > profiles/rNN_energy_by_class.json
Reports, per class: time per call, average board power, J per call, SM clock; and the same for the whole step, next to the
sum over classes weighted by the calls one step makes."""
import json
import os
import sys
import threading
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pynvml
import torch

from anyref_b200 import ops
from anyref_b200.grounding import GroundingPath
from anyref_b200.segment_anything import build_sam_from_config
from anyref_b200.synthetic import CONFIGS, synthetic_images, synthetic_seg_embeddings, synthetic_state_dict

dt = torch.bfloat16 if (len(sys.argv) > 1 and sys.argv[1] == "bf16") else torch.float16
dev = torch.device("cuda:0")
pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
B, heads, E, Hm = 16, 16, 1280, 5120
M = B * 4096
torch.manual_seed(0)


def measure(fn, seconds=2.0):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    fn()
    torch.cuda.synchronize()
    per = max(time.perf_counter() - t0, 1e-5)
    iters = max(5, int(seconds / per))
    clocks = []
    stop = threading.Event()

    def sampler():
        while not stop.is_set():
            clocks.append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
            time.sleep(0.05)

    th = threading.Thread(target=sampler)
    th.start()
    torch.cuda.synchronize()
    e0 = pynvml.nvmlDeviceGetTotalEnergyConsumption(h)
    t0 = time.perf_counter()
    for _ in range(iters):
        fn()
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    e1 = pynvml.nvmlDeviceGetTotalEnergyConsumption(h)
    stop.set()
    th.join()
    sec = t1 - t0
    clocks.sort()
    return {"ms_per_call": sec / iters * 1e3, "watts": (e1 - e0) * 1e-3 / sec, "joules_per_call": (e1 - e0) * 1e-3 / iters,
            "sm_mhz_median": clocks[len(clocks) // 2] if clocks else None, "calls": iters}


def rand(*shape, scale=1.0):
    return (torch.randn(*shape, device=dev) * scale).to(dt)


out = {"dtype": str(dt), "how": "each class back to back for ~2 s, NVML energy counter around the loop; B = 16, ViT-H shapes",
       "classes": {}}
x = rand(M, E)
# the four encoder linears with their production epilogues (plain variants: no LayerNorm folding, which needs its stats inputs)
w_qkv, w_proj, w_l1, w_l2 = rand(3 * E, E, scale=0.02), rand(E, E, scale=0.02), rand(Hm, E, scale=0.02), rand(E, Hm, scale=0.02)
res = torch.randn(M, E, device=dev)
hbuf = rand(M, Hm)
out["classes"]["gemm_qkv"] = dict(measure(lambda: ops.gemm(x, w_qkv, bias=None, out_dtype=dt)), calls_per_step=32)
out["classes"]["gemm_proj_residual"] = dict(measure(lambda: ops.gemm(x, w_proj, residual=res, out=res)), calls_per_step=32)
out["classes"]["gemm_lin1_gelu"] = dict(measure(lambda: ops.gemm(x, w_l1, act="gelu", out_dtype=dt)), calls_per_step=32)
out["classes"]["gemm_lin2_residual"] = dict(measure(lambda: ops.gemm(hbuf, w_l2, residual=res, out=res)), calls_per_step=32)
del hbuf, res
qkv = rand(M, 3 * E)
bias = rand(3 * E)
tab = ops.window_rel_table(torch.randn(27, 80, device=dev) * 0.1, torch.randn(27, 80, device=dev) * 0.1, dt)
gh = ops.global_rel_table(torch.randn(127, 80, device=dev) * 0.1, dt)
gw = ops.global_rel_table(torch.randn(127, 80, device=dev) * 0.1, dt)
out["classes"]["attn_window"] = dict(measure(lambda: ops.attn_window(qkv, bias, tab, B, heads)), calls_per_step=28)
out["classes"]["attn_global"] = dict(measure(lambda: ops.attn_global(qkv, gh, gw, B, heads)), calls_per_step=4)
del qkv, x
torch.cuda.empty_cache()

# the whole step
cfg = CONFIGS["vit_h"]
sam = build_sam_from_config(cfg)
sam.load_state_dict(synthetic_state_dict(cfg, seed=1234))
sam = sam.to(dev).eval()
sam.image_encoder.set_operand_dtype(dt)
path = GroundingPath(sam)
images = synthetic_images(B, seed=0).to(dev)
seg = [s.to(dev) for s in synthetic_seg_embeddings(B, 1, seed=0)]
sizes = [(1024, 1024)] * B
with torch.no_grad():
    out["step"] = measure(lambda: path(images, seg, sizes, sizes), seconds=4.0)
parts = sum(c["joules_per_call"] * c["calls_per_step"] for c in out["classes"].values())
tparts = sum(c["ms_per_call"] * c["calls_per_step"] for c in out["classes"].values())
out["sum_of_classes"] = {"joules_per_step": parts, "ms_per_step": tparts,
                         "note": "classes timed alone run at a higher clock than inside the step; the J per call is the comparable "
                                 "quantity under a power cap"}
idle0 = pynvml.nvmlDeviceGetTotalEnergyConsumption(h)
time.sleep(1.0)
out["idle_watts"] = (pynvml.nvmlDeviceGetTotalEnergyConsumption(h) - idle0) * 1e-3
out["power_limit_w"] = pynvml.nvmlDeviceGetEnforcedPowerLimit(h) * 1e-3
print(json.dumps(out, indent=1))
