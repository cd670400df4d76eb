"""Where a training step of the mask decoder spends its time: host phases (perf_counter, synchronised) at n prompts."""
import ctypes as C
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from anyref_b200 import _lib
from anyref_b200.segment_anything import _pack, build_sam_from_config
from anyref_b200.synthetic import CONFIGS, synthetic_state_dict

cfg = CONFIGS["vit_tiny80"]
sam = build_sam_from_config(cfg)
sam.load_state_dict(synthetic_state_dict(cfg))
sam = sam.cuda()
dec = sam.mask_decoder
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1
lib = _lib.load()
emb = torch.randn(1, 256, 64, 64, device="cuda") * 0.5
pe = sam.prompt_encoder.get_dense_pe().float()
sparse = torch.randn(n, 1, 256, device="cuda")
dense_vec = sam.prompt_encoder.no_mask_embed.weight.detach().reshape(-1).contiguous()


def sync():
    torch.cuda.synchronize()
    return time.perf_counter()


res = {}
for it in range(6):
    t0 = sync()
    shape, blob = _pack.pack_decoder(dec, 64)
    t1 = sync()
    nbytes = lib.sam_decoder_train_workspace_bytes(C.byref(shape), n, 1)
    ws = torch.empty(nbytes + 256, dtype=torch.uint8, device="cuda")
    masks = torch.empty(n, 4, 256, 256, device="cuda")
    iou = torch.empty(n, 4, device="cuda")
    tape = C.c_void_p()
    t2 = sync()
    l0 = lib.sam_launch_count()
    rc = lib.sam_decoder_train_forward(C.byref(shape), blob.data_ptr(), emb.data_ptr(), 2, 1, None, sparse.data_ptr(), n, 1,
                                       dense_vec.data_ptr(), None, 2, pe.data_ptr(), 2, masks.data_ptr(), iou.data_ptr(),
                                       ws.data_ptr(), ws.numel(), C.byref(tape), None)
    assert rc == 0, lib.sam_last_error()
    t3h = time.perf_counter()
    t3 = sync()
    l1 = lib.sam_launch_count()
    dm = torch.randn_like(masks)
    gblob = torch.zeros_like(blob)
    ds = torch.empty_like(sparse)
    t4 = sync()
    rc = lib.sam_decoder_backward(tape, dm.data_ptr(), 0, 1, None, gblob.data_ptr(), ds.data_ptr(), None)
    assert rc == 0, lib.sam_last_error()
    t5h = time.perf_counter()
    t5 = sync()
    l2 = lib.sam_launch_count()
    lib.sam_decoder_tape_free(tape)
    grads = _pack.unpack_decoder_grads(dec, gblob)
    t6 = sync()
    res = {"pack": t1 - t0, "alloc": t2 - t1, "forward host": t3h - t2, "forward total": t3 - t2, "backward host": t5h - t4,
           "backward total": t5 - t4, "unpack": t6 - t5, "launches fwd": l1 - l0, "launches bwd": l2 - l1,
           "workspace MB": nbytes / 2**20}
print(f"n={n}: " + ", ".join(f"{k} {v * 1e3:.2f} ms" if isinstance(v, float) and k != "workspace MB" else f"{k} {v:.0f}" for k, v in res.items()))
