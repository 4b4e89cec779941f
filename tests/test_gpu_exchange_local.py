"""The N > 1 exchange of the sharded MSM on a ONE-GPU box: R contexts of one process act as the ranks
(zc_peer_mailbox_connect_local), so the driver's single-GPU test run covers the peer-store / flag / fold kernel against the
oracle (the two-GPU test is skipped there).  Also the two robustness holes of round 1: back-to-back tiny
exchanges (double-buffered slots) and a rank that never arrives (bounded wait -> ZC_ERR_STATE)."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORKER = os.path.join(ROOT, "tests", "exchange_local_worker.py")


def _run(args, extra_env=None, timeout=300):
    env = dict(os.environ, CUDA_DEVICE_MAX_CONNECTIONS="32")
    env.update(extra_env or {})
    p = subprocess.run([sys.executable, WORKER] + [str(a) for a in args], capture_output=True, text=True, env=env, timeout=timeout)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-4000:]
    return json.loads(p.stdout.strip().splitlines()[-1])


@pytest.mark.parametrize("ranks,calls,n", [(2, 1000, 64), (4, 200, 64), (8, 40, 3000), (3, 60, 1000)])
def test_sharded_exchange_on_one_gpu_vs_oracle(ranks, calls, n):
    r = _run(["stress", ranks, calls, n])
    assert r["rank_bit_mismatches"] == 0, r
    assert r["oracle_mismatches"] == 0, r
    assert r["generators_path_ok"], r
    assert r["launches"] > 0


def test_missing_rank_is_an_error_not_a_hang():
    r = _run(["timeout"], {"ZC_PEER_TIMEOUT_MS": "300"}, timeout=120)
    assert r["status"] == 5, r                      # ZC_ERR_STATE
    assert "rank 1" in r["message"], r
    assert r["identity_returned"], r
    assert r["ok_after_reconnect"], r
