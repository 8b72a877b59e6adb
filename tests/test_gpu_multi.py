"""Multi-GPU path on real peers (needs >= 2 GPUs on the box; skipped otherwise): two ranks under
torchrun shard one pair list in interleaved blocks, the alignment kernel's epilogue stores every record
into BOTH ranks' symmetric-memory buffers over NVLink, and bench.py compares every row of every
rank's buffer with an NCCL all_gather of the same records (buffers zeroed first)."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:                                         # noqa: BLE001
        return 0


@pytest.mark.skipif(_gpus() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("workload,extra", [("proximity", ["--scans", "1500", "--beams", "128", "--pairs", "3000", "--block", "64"]),
                                            ("allpairs", ["--scans", "80", "--beams", "360", "--block", "100"]),
                                            ("chain", ["--scans", "400", "--beams", "360"])])
def test_fused_gather_across_two_gpus(workload, extra):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29731", os.path.join(ROOT, "bench.py"), "--gpus", "2", "--steps", "2",
           "--warmup", "1", "--workload", workload, "--no-cpu", "--no-sustained", "--no-exhaustive"] + extra
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-3000:]
    line = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    assert line["n_gpus"] == 2 and line["details"]["gather_check"].startswith("fused gather == NCCL all_gather")
    assert line["scaling"] == ("weak" if workload == "chain" else "strong")
    assert line["e2e"]["value"] > 0 and line["gpu_launches"] == 2
