"""CPU: the loop-level fixtures load and the tiny victim reproduces the label the reference attack saw."""
import numpy as np
import torch

import tiny_victim
from conftest import load_golden


def test_tiny_victim_fixture_round_trip():
    g = load_golden("l4_attack_loops")
    v = tiny_victim.from_npz(g)
    with torch.no_grad():
        label = v(torch.from_numpy(g["data"]).transpose(1, 2))[0].argmax(1)
    assert np.array_equal(label.numpy(), g["label"])
    assert g["cw_chamfer_noise"].shape == (3, 1, 3, 256) and g["knn_noise"].shape == (1, 1, 3, 256)
    # the reference attacks moved the cloud and stayed inside the clip budgets
    assert 1e-3 < np.abs(g["cw_chamfer_bestattack"] - g["data"]).max() <= 0.18 + 1e-6
    assert 1e-3 < np.linalg.norm(g["knn_adv"] - g["data"], axis=-1).max() <= 0.1 + 1e-6
