"""CPU checks of the dropout-mask restatement (oracle.philox_keep <-> csrc/hh_rng.cuh): distribution, scaling,
determinism and independence of sites / steps.  The GPU tests (tests/test_gpu_backward.py) check that the CUDA kernels
draw exactly these masks."""
import math

import torch

from oracle import hh_oracle as O


def test_keep_rate_and_scale():
    n = 400000
    for p in (0.1, 0.3):
        m = O.philox_keep(12345, 0, 3, n, p)
        vals = set(m.unique().tolist())
        assert vals == {0.0, float(torch.tensor(1.0 / (1.0 - p), dtype=torch.float32))}
        keep = float((m > 0).float().mean())
        sigma = math.sqrt(p * (1 - p) / n)
        assert abs(keep - (1 - p)) < 5 * sigma + 1e-4          # 16-bit threshold: |p_eff - p| < 2^-16
        assert abs(float(m.mean()) - 1.0) < 6 * sigma / (1 - p) + 1e-3   # E[mask] = 1, as for F.dropout


def test_streams_are_reproducible_and_distinct():
    a = O.philox_keep(7, 2, 9, 4096, 0.1)
    assert torch.equal(a, O.philox_keep(7, 2, 9, 4096, 0.1))
    for other in (O.philox_keep(8, 2, 9, 4096, 0.1), O.philox_keep(7, 3, 9, 4096, 0.1), O.philox_keep(7, 2, 10, 4096, 0.1)):
        agree = float((a == other).float().mean())
        assert 0.75 < agree < 0.9                               # independent masks agree on ~ p^2 + (1-p)^2 = 0.82


def test_known_answer_philox():
    """Random123 known-answer vectors of Philox4x32-10 (counter = key = 0 and the all-ones vector): the generator behind
    the masks is the standard one."""
    import numpy as np

    def philox(c, k):
        c = [np.uint64(x) for x in c]
        k = [np.uint64(x) for x in k]
        M = np.uint64(0xFFFFFFFF)
        for _ in range(10):
            p0 = np.uint64(0xD2511F53) * c[0]
            p1 = np.uint64(0xCD9E8D57) * c[2]
            c = [((p1 >> np.uint64(32)) ^ c[1] ^ k[0]) & M, p1 & M, ((p0 >> np.uint64(32)) ^ c[3] ^ k[1]) & M, p0 & M]
            k = [(k[0] + np.uint64(0x9E3779B9)) & M, (k[1] + np.uint64(0xBB67AE85)) & M]
        return [int(x) for x in c]
    assert philox([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert philox([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]


def test_eval_mode_is_identity():
    x = torch.randn(3, 5)
    assert O._drop(x, None, 0) is x
    assert O._drop(x, {"p": 0.0, "seed": 1, "offset": 0}, 0) is x
