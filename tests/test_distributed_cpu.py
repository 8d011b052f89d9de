"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: the flat-gradient all-reduce that makes
data-parallel training one collective per step, and bench.py's rank handling under torchrun-style env."""
import json
import os
import socket
import subprocess
import sys
from pathlib import Path

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent


def _free_port() -> int:
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank: int, world: int, port: int, out_dir: str) -> None:
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, str(ROOT))
    from fastervit_b200.engine_train import allreduce_mean_
    import fastervit_b200 as F
    # identical replicas (same seed), rank-dependent "gradients"
    torch.manual_seed(0)
    model = F.create_model("faster_vit_0_224", dim=16, in_dim=16, depths=[1, 1, 1, 1], num_heads=[1, 2, 4, 8])
    n = sum(p.numel() for p in model.parameters())
    flat = torch.arange(n, dtype=torch.float32) * (rank + 1)
    allreduce_mean_(flat)
    expect = torch.arange(n, dtype=torch.float32) * (sum(range(1, world + 1)) / world)
    ok = torch.allclose(flat, expect)
    # the bucketed reducer used by the backward launch list: slices issued out of order + a remainder
    from fastervit_b200.engine_train import GradBucketReducer
    flat2 = torch.arange(n, dtype=torch.float32) * (rank + 1)
    cuts = [0, n // 5, n // 2, n]
    buckets = [(cuts[2], cuts[3]), (cuts[1], cuts[2]), (cuts[0], cuts[1])]   # completion order of the backward
    red = GradBucketReducer(flat2)
    red.reduce(*buckets[0])
    red.reduce(*buckets[1])
    red.finish(buckets)
    ok = ok and torch.allclose(flat2, expect)
    # parameters of both replicas are bit-identical (what the gradient-only exchange relies on)
    digest = float(sum(p.double().sum() for p in model.parameters()))
    gathered = [None] * world
    dist.all_gather_object(gathered, digest)
    Path(out_dir, f"rank{rank}.json").write_text(json.dumps({"ok": bool(ok), "same_init": len(set(gathered)) == 1}))
    dist.destroy_process_group()


def test_flat_gradient_allreduce_gloo_world2(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        res = json.loads((tmp_path / f"rank{r}.json").read_text())
        assert res["ok"] and res["same_init"], res


def test_allreduce_requires_process_group():
    from fastervit_b200.engine_train import allreduce_mean_
    from fastervit_b200.lib import FvitError
    if dist.is_initialized():
        pytest.skip("a process group is already initialised in this interpreter")
    with pytest.raises(FvitError):
        allreduce_mean_(torch.zeros(4))


def test_bench_reference_arm_only_rank0_prints():
    env = dict(os.environ, WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(_free_port()))
    outs = []
    for rank in (1, 0):
        env["RANK"] = env["LOCAL_RANK"] = str(rank)
        r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                            "--warmup", "1"], env=env, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-500:]
        outs.append(r.stdout.strip())
    assert outs[0] == ""                       # rank 1: exits 0 without work
    line = json.loads(outs[1].splitlines()[-1])  # rank 0: one JSON line
    assert line["impl"] == "reference" and line["n_gpus"] == 2 and line["e2e"]["h2d_bytes_per_step"] == 0


def test_bench_clock_sampler_windows_samples_to_the_timed_region(tmp_path, monkeypatch):
    """bench.py's nvidia-smi sampler: lines are time-stamped on arrival, only those inside [mark_begin, mark_end] count,
    throttle reasons are collected, and a timed region too short for two samples falls back to warm-up + timed."""
    import importlib.util
    import os
    import stat
    import time
    fake = tmp_path / "nvidia-smi"
    fake.write_text("#!/bin/bash\nsleep 0.3\ni=0\nwhile true; do\n"
                    "if [ $((i % 2)) -eq 0 ]; then echo '1935, 1965, 998.1, Not Active, Not Active, Not Active, Active'; "
                    "else echo '1965, 1965, 800.0, Not Active, Not Active, Not Active, Not Active'; fi\n"
                    "i=$((i+1)); sleep 0.05\ndone\n")
    fake.chmod(fake.stat().st_mode | stat.S_IEXEC)
    monkeypatch.setenv("PATH", f"{tmp_path}{os.pathsep}{os.environ['PATH']}")
    spec = importlib.util.spec_from_file_location("bench_mod", ROOT / "bench.py")
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    s = bench.ClockSampler(0)
    s.start()
    time.sleep(0.8)            # "warm-up": samples arrive but must not be counted
    s.mark_begin()
    time.sleep(0.5)
    s.mark_end()
    time.sleep(0.2)            # samples after the region must not be counted either
    out = s.stop()
    assert out["window"] == "timed region" and 3 <= out["samples"] <= 40
    assert out["sm_max_mhz"] == 1965.0 and out["sm_mhz"] in (1935.0, 1965.0) and out["reasons"] == ["sw_power_cap"]
    s = bench.ClockSampler(0)
    s.start()
    time.sleep(0.7)
    s.mark_begin()
    time.sleep(0.01)
    s.mark_end()
    out = s.stop()
    assert out["window"].startswith("warm-up") and out["samples"] >= 2
