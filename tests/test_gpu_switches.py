"""The library's A/B switches change scheduling (programmatic dependent launch, the side branch for weight gradients,
the side stream for step preparation, stored vs re-hashed dropout decisions, tcgen05 vs mma.sync attention), never results: one training step under each
switch, in a fresh process (the switches are read once per process), against the default build."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env, *args):
    e = dict(os.environ)
    e.update(env)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "step_checksum.py"), *args], cwd=ROOT, env=e,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    return json.loads(r.stdout.strip().splitlines()[-1])


def _close(a, b, tol):
    if isinstance(a, list):
        return all(_close(x, y, tol) for x, y in zip(a, b))
    return abs(a - b) <= tol * max(abs(a), abs(b), 1e-12)


@pytest.mark.parametrize("shape", ["short", "long"])
def test_switches_do_not_change_results(shape):
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    args = ("long",) if shape == "long" else ()
    ref = _run({}, *args)
    # scheduling-only switches: same kernels, same arithmetic (fp32 atomics may land in another order: 1e-5)
    for env in ({"SER_PDL": "0"}, {"SER_SIDE_STREAM": "0"}, {"SER_PDL": "0", "SER_SIDE_STREAM": "0"}):
        got = _run(env, *args)
        assert set(got) == set(ref)
        for k in ref:
            assert _close(got[k], ref[k], 1e-5), (env, k, got[k], ref[k])
    # step preparation (classifier weight cast, zero fill of the persistent gradient arena) beside the fusion chain on a
    # side stream vs at the start of the step: second of two steps on one persistent arena
    ref_p = _run({}, *args, "persist")
    got = _run({"SER_PREP_SIDE": "0"}, *args, "persist")
    for k in ref_p:
        assert _close(got[k], ref_p[k], 1e-5), ("prep side", k, got[k], ref_p[k])
    # stored vs re-hashed dropout decisions: the same masks, one fused multiply-add contracts differently (bf16 last places)
    got = _run({"SER_ATTN_KEEPBITS": "0"}, *args)
    for k in ref:
        assert _close(got[k], ref[k], 2e-3), ("keepbits", k, got[k], ref[k])
    # fused attention backward (one kernel per small problem) vs the dQ + dK/dV pair: the same arithmetic per element,
    # dQ summed over the keys in another order (bf16 last places)
    got = _run({"SER_ATTN_BWD_FUSED": "0"}, *args)
    for k in ref:
        assert _close(got[k], ref[k], 2e-3), ("fused attention backward", k, got[k], ref[k])
    # tcgen05 vs mma.sync attention kernels: two implementations of the same maths in bf16 -- forward quantities agree at
    # the bf16 tolerance; gradients only at the bf16 floor of this head (ReLU-gate flips)
    got = _run({"SER_ATTN_FWD": "1", "SER_ATTN_BWD": "1"}, *args)
    for k in ref:
        assert _close(got[k], ref[k], 2e-2 if k in ("loss", "logits") else 0.25), ("mma.sync", k, got[k], ref[k])
