"""CPU tests: the oracle restatement vs the committed golden vectors that oracle/make_golden.py
produced from the unmodified reference (fp64). This is what pins the oracle."""
import pytest
import torch

from oracle import fastervit_oracle as O
from oracle.configs import CASES, cfg_of

from pathlib import Path

GOLDEN = Path(__file__).parent / "golden"


def _load(case):
    return torch.load(GOLDEN / f"{case}.pt", weights_only=False)


def _state_dict(g, dtype):
    sd = {}
    for k, shape, dt in g["state_keys"]:
        sd[k] = torch.zeros(shape, dtype=dtype if dt.startswith("float") else getattr(torch, dt))
    # deterministic buffers the reference builds in __init__ (fv.py:226-254)
    for k in list(sd):
        if k.endswith("relative_coords_table"):
            ws = (sd[k].shape[1] + 1) // 2
            sd[k] = O.rel_coords_table(ws, dtype)
        elif k.endswith("relative_position_index"):
            ws = int(round(sd[k].shape[0] ** 0.5))
            sd[k] = O.rel_position_index(ws)
    return O.synth_fill_(sd, g["seeds"]["w"])


def _sample(t, n=512):
    f = t.detach().flatten()
    stride = max(1, (f.numel() + n - 1) // n)
    return f[::stride].float()


@pytest.mark.parametrize("case", ["tiny_a", "tiny_qk", "tiny_b", "tiny_ar", "tiny_ar85", "tiny_ar68", "tiny_ar148", "tiny_21k", "tiny_21k224", "fv0", "fv4",
                                  "ar0"])
def test_oracle_eval_matches_reference_fp64(case):
    g = _load(case)
    sd = _state_dict(g, torch.float64)
    x = O.synth_input(g["eval"]["batch"], g["cfg"]["resolution"], g["seeds"]["x"], torch.float64)
    cap = {}
    with torch.no_grad():
        logits = O.forward(sd, g["cfg"], x, capture=cap)
    ref = g["eval"]["logits"]
    assert logits.shape == ref.shape
    err = ((logits - ref).abs().max() / ref.abs().max()).item()
    assert err < 1e-9, err
    # per-level activations against the reference's forward-hook samples
    for i in range(4):
        key = f"levels.{i}"
        got = cap[f"levels.{i}.down"] if i < 3 else cap[f"levels.{i}.out"]
        want = g["eval"]["acts"][key]
        assert tuple(got.shape) == tuple(want["shape"])
        d = (_sample(got) - want["sample"]).abs().max() / want["amax"]
        assert d < 1e-5, (key, d)


@pytest.mark.parametrize("case", ["tiny_a", "tiny_qk", "tiny_b", "tiny_ar", "tiny_ar85", "tiny_ar68", "tiny_ar148", "tiny_21k", "tiny_21k224"])
def test_oracle_train_grads_match_reference(case):
    g = _load(case)
    sd = _state_dict(g, torch.float64)
    tr = g["train"]
    x = O.synth_input(tr["batch"], g["cfg"]["resolution"], g["seeds"]["x"] + 100, torch.float64)
    loss, logits, grads = O.loss_and_grads(sd, g["cfg"], x, tr["target"], training=True)
    assert abs(loss.item() - tr["loss"]) < 1e-9 * max(1.0, abs(tr["loss"]))
    assert ((logits - tr["logits"]).abs().max() / tr["logits"].abs().max()).item() < 1e-9
    floor = tr["grad_floor"]
    assert set(grads) == set(tr["grads"])
    for k, want in tr["grads"].items():
        got = grads[k]
        assert tuple(got.shape) == tuple(want["shape"])
        assert abs(got.norm().item() - want["l2"]) <= 1e-6 * max(want["l2"], floor), k
        d = (_sample(got) - want["sample"]).abs().max().item()
        assert d <= 2e-6 * max(want["amax"], floor), (k, d)


def test_oracle_fp32_close_to_fp64_golden():
    g = _load("fv0")
    sd = _state_dict(g, torch.float32)
    x = O.synth_input(g["eval"]["batch"], 224, g["seeds"]["x"], torch.float32)
    with torch.no_grad():
        logits = O.forward(sd, g["cfg"], x)
    ref = g["eval"]["logits"]
    assert ((logits.double() - ref).abs().max() / ref.abs().max()).item() < 2e-5


def test_fp16_operand_model_error_budget():
    """The fp16-operand arithmetic model (what the CUDA path computes) stays inside the 1e-3 target."""
    g = _load("fv0")
    sd = _state_dict(g, torch.float32)
    x = O.synth_input(g["eval"]["batch"], 224, g["seeds"]["x"], torch.float32)
    with torch.no_grad():
        logits = O.forward(sd, g["cfg"], x, quant="fp16")
    ref = g["eval"]["logits"]
    err = ((logits.double() - ref).abs().max() / ref.abs().max()).item()
    print("fp16-operand model error on fv0:", err)
    assert err < 1e-3


def test_golden_schema_has_reference_param_counts():
    assert _load("fv0")["n_params"] == 31404840   # README.md:152 (31.4 M)
    assert _load("fv4")["n_params"] == 365555712
