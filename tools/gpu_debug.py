"""Developer tool (GPU box): run each low-level check in its own subprocess so that a trapped or
hung kernel cannot poison the others; append one JSON line per check to gpurun_out/debug.jsonl.

    python tools/gpu_debug.py [--only selftest|attn] [--timeout 60]
"""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
os.makedirs(OUT, exist_ok=True)

CHILD = r'''
import json, sys, os
ROOT = {root!r}
sys.path[:0] = [os.path.join(ROOT, "flash-attention-v2-rdna3-minimal_b200"), os.path.join(ROOT, "oracle")]
import torch
from rocwmma_fattn import _capi
import fa_oracle as orc
spec = json.loads({spec!r})
res = dict(spec)
try:
    if spec["kind"] == "selftest":
        dt = torch.float16 if spec["dtype"] == "f16" else torch.bfloat16
        g = torch.Generator().manual_seed(1)
        a = torch.randn(128, 128, generator=g).to(dt).cuda()
        b = torch.randn(128, 128, generator=g).to(dt).cuda()
        out = torch.full((128, 128), float("nan"), device="cuda")
        rc = _capi.lib.fa_umma_selftest(a.data_ptr(), b.data_ptr(), out.data_ptr(), 0 if dt == torch.float16 else 1,
                                        spec["mode"], spec["lbo"], spec["sbo"], None)
        res["rc"] = rc
        torch.cuda.synchronize()
        ref = a.float() @ (b.float().t() if spec["mode"] == 0 else b.float())
        res["max_err"] = (out - ref).abs().max().item()
        res["ref_max"] = ref.abs().max().item()
        res["nan"] = int(torch.isnan(out).sum().item())
    else:
        from rocwmma_fattn.FlashAttn import flash_attn_forward
        dt = torch.float16 if spec["dtype"] == "f16" else torch.bfloat16
        B, H, Nq, Nkv, D = spec["shape"]
        q, k, v = orc.make_inputs(B, H, Nq, Nkv, D, dt, seed=3, dist=spec.get("dist", "rand"))
        ref, lse_ref = orc.sdpa_math(q, k, v, causal=spec["causal"])
        _capi.set_kernel(spec["kernel"])
        o, lse = flash_attn_forward(q.cuda(), k.cuda(), v.cuda(), spec["causal"], return_lse=True)
        torch.cuda.synchronize()
        res["max_err"] = orc.max_abs_err(o, ref)
        res["lse_err"] = (lse.double().cpu() - lse_ref.double()).abs().max().item()
        res["nan"] = int((~torch.isfinite(o.float())).sum().item())
        # where are the errors? per 128-row tile / per 64-col block summary
        e = (o.float().cpu() - ref).abs()
        res["err_by_rowtile"] = [round(x, 5) for x in e.amax(dim=(0, 1, 3)).reshape(-1)[::max(1, Nq // 8)].tolist()][:8]
        res["err_by_colblk"] = [round(e[..., c:c + 32].max().item(), 5) for c in range(0, D, 32)]
except Exception as ex:
    res["error"] = repr(ex)[:500]
print("RESULT " + json.dumps(res))
'''


def run_child(spec, timeout):
    code = CHILD.format(root=ROOT, spec=json.dumps(spec))
    t0 = time.time()
    try:
        p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=timeout)
        out = p.stdout
        res = None
        for ln in out.splitlines():
            if ln.startswith("RESULT "):
                res = json.loads(ln[7:])
        if res is None:
            res = dict(spec, error="no result", rc_proc=p.returncode, stderr=p.stderr[-800:], stdout=out[-300:])
    except subprocess.TimeoutExpired:
        res = dict(spec, error="timeout")
    res["secs"] = round(time.time() - t0, 1)
    with open(os.path.join(OUT, "debug.jsonl"), "a") as fh:
        fh.write(json.dumps(res) + "\n")
    print(json.dumps(res), flush=True)
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="")
    ap.add_argument("--timeout", type=int, default=90)
    a = ap.parse_args()
    ok_modes = {}
    if a.only in ("", "selftest"):
        for mode in (0, 1, 2, 3):
            r = run_child(dict(kind="selftest", dtype="f16", mode=mode, lbo=0, sbo=0), a.timeout)
            good = "error" not in r and r.get("max_err", 1e9) < 0.05
            ok_modes[mode] = good
            if not good and mode in (1, 2):
                # descriptor search for the MN-major B operand
                for lbo, sbo in ((1024, 16384), (16384, 128), (128, 16384), (2048, 1024), (1024, 2048), (16, 1024)):
                    r2 = run_child(dict(kind="selftest", dtype="f16", mode=mode, lbo=lbo, sbo=sbo), a.timeout)
                    if "error" not in r2 and r2.get("max_err", 1e9) < 0.05:
                        break
        run_child(dict(kind="selftest", dtype="bf16", mode=0, lbo=0, sbo=0), a.timeout)
    if a.only in ("", "attn"):
        shapes = [((1, 1, 128, 128, 64), False), ((1, 1, 128, 128, 128), False), ((1, 2, 256, 256, 128), False),
                  ((1, 2, 512, 512, 128), True), ((1, 2, 1000, 900, 128), False), ((2, 2, 200, 77, 64), True)]
        for kernel in (1, 2, 3, 4):
            for shape, causal in shapes:
                run_child(dict(kind="attn", dtype="f16", kernel=kernel, shape=shape, causal=causal), a.timeout)
        run_child(dict(kind="attn", dtype="bf16", kernel=4, shape=(1, 16, 2048, 2048, 128), causal=False, dist="randn"), a.timeout)
        run_child(dict(kind="attn", dtype="f16", kernel=4, shape=(1, 16, 2048, 2048, 128), causal=True, dist="randn"), a.timeout)


if __name__ == "__main__":
    main()
