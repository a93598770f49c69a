"""CPU: the host half of get_stationary_distribution -- the reference's recurrent-class rule
(colosseum/mdp/utils/markov_chain.py:90-133) restated in colosseum_b200.markov_chain.recurrent_class_weights --
against vectors recorded from the reference (tests/golden/multichain.npz, avg_reward.npz); the limit itself is taken
by the oracle here and by the GPU in tests/test_gpu_markov.py."""
import os

import numpy as np

from conftest import GOLDEN
from oracle import oracle as orc


def test_recurrent_class_rule_matches_the_reference():
    import colosseum_b200.markov_chain as mc

    g = np.load(os.path.join(GOLDEN, "multichain.npz"))
    for name in g["names"]:
        starts = list(zip(g[f"{name}_start_idx"].tolist(), g[f"{name}_start_prob"].tolist()))
        x0, n = mc.recurrent_class_weights(g[f"{name}_tps"], starts)
        assert n > 1 and abs(x0.sum() - 1) < 1e-12
        sd = orc.stationary_distribution_f64(g[f"{name}_tps"].astype(np.float32), x0)
        np.testing.assert_allclose(sd, g[f"{name}_sd"], atol=1e-7, err_msg=str(name))


def test_unichain_policies_ignore_the_start_distribution():
    import colosseum_b200.markov_chain as mc

    g = np.load(os.path.join(GOLDEN, "avg_reward.npz"))
    for name in g["names"]:
        for k in ("opt", "worst", "rand"):
            tps = g[f"{name}_{k}_tps"]
            x0, n = mc.recurrent_class_weights(tps, None)  # one class: the start distribution is not needed
            assert n == 1 and x0.sum() == 1.0
            sd = orc.stationary_distribution_f64(tps, x0)
            np.testing.assert_allclose(sd, g[f"{name}_{k}_sd"], atol=2e-5, err_msg=f"{name} {k}")
