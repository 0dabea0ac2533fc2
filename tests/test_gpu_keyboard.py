"""Word prototypes / minimum-jerk trajectories on the GPU against fixtures produced by the
unmodified reference's QWERTYKeyboard (oracle/make_keyboard_golden.py -> tests/golden/keyboard_golden.npz): 100 words
incl. one-key, repeated-key, no-key and 26-key words.  Tolerance: 2e-6 absolute on the float32 (x, y, t) rows (float64
arithmetic on both sides; the GPU contracts multiply-adds)."""
import os

import numpy as np
import pytest
import torch

from wgg_b200.keyboard_gpu import QWERTYKeyboard

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "keyboard_golden.npz"))
WORDS = [str(w) for w in G["words"]]
TOL = 2e-6


def test_key_centers_match_reference():
    kb = QWERTYKeyboard(device="cpu")
    mine = np.array([kb.key_centers[c] for c in "abcdefghijklmnopqrstuvwxyz"])
    assert np.array_equal(mine, G["key_centers"])


@pytest.mark.gpu
@pytest.mark.parametrize("T", [128, 37])
def test_word_prototypes(T):
    kb = QWERTYKeyboard(device="cuda:0")
    out = kb.get_word_prototypes(WORDS, T).cpu().numpy()
    ref = G[f"proto_T{T}"]
    assert out.shape == ref.shape and out.dtype == np.float32
    err = np.abs(out.astype(np.float64) - ref).max(axis=(1, 2))
    assert err.max() <= TOL, (WORDS[int(err.argmax())], err.max())
    assert kb.get_word_prototypes([], T).shape == (0, T, 3)


@pytest.mark.gpu
@pytest.mark.parametrize("key,T,mid,std,seed", [("mj_mid_clean", 128, True, 0.0, None), ("mj_nomid_clean", 128, False, 0.0, None),
                                                ("mj_mid_noise003", 128, True, 0.03, 5), ("mj_nomid_noise005_T64", 64, False, 0.05, 6)])
def test_minimum_jerk(key, T, mid, std, seed):
    kb = QWERTYKeyboard(device="cuda:0")
    if seed is not None:
        np.random.seed(seed)   # the wrapper draws the offsets from numpy's global generator in the reference's order
    out = kb.get_minimum_jerk_trajectories(WORDS, T, include_midpoints=mid, offset_std=std).cpu().numpy()
    ref = G[key]
    err = np.abs(out.astype(np.float64) - ref).max(axis=(1, 2))
    assert err.max() <= TOL, (WORDS[int(err.argmax())], err.max())
    # time is monotone and spans [0, 1] for every word with at least two distinct keys
    multi = [i for i, w in enumerate(WORDS) if len(set(c for c in w if c.isalpha())) >= 2]
    assert (np.diff(out[multi, :, 2], axis=1) >= -1e-7).all() and np.allclose(out[multi, -1, 2], 1.0)
