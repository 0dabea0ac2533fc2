"""Realistic (prototype, gesture) fixture for the acceptance run (SURVEY.md 8(d) "realistic set", VERDICT N2).

Run in the build container only (needs /root/reference):  python -m oracle.make_realistic_fixture

The dataset of the reference (swipelogs.zip) is not in the checkout, so gestures are synthesised with the reference's
OWN code, unmodified: words from dataset/wordfreq.txt (alphabetic, length >= 2, data.py:201), prototype =
QWERTYKeyboard().get_word_prototype(word, 128) (keyboard.py:710-765), "real" gesture =
generate_minimum_jerk_trajectory(key centres, 128, include_midpoints=True, offset_std ~ U(0.02, 0.05))
(keyboard.py:389-514) with np.random.seed(0); up to 5 gestures per word, 80/20 WORD-level split (config.py:61-62,
data.py:444-505); values clipped like data.py:413.  Stored as float16 (the coordinates live in [-1, 1]; 5e-4 absolute
is far below the 0.02 key-offset noise) in tests/golden/realistic_gestures.npz together with the words.
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle.ref_loader import load_reference  # noqa: E402

N_WORDS, PER_WORD, T = 1500, 3, 128


def main():
    ref = load_reference()
    kb = ref.keyboard.QWERTYKeyboard()
    words = []
    for line in open(os.path.join(ref.root, "dataset", "wordfreq.txt")):
        parts = line.split()
        if len(parts) == 2 and parts[1].isalpha() and parts[1].isascii() and len(parts[1]) >= 2:
            words.append((int(parts[0]), parts[1].lower()))
    words.sort(key=lambda w: (-w[0], w[1]))          # most frequent first, deterministic
    words = [w for _, w in words[:N_WORDS]]
    rng = np.random.RandomState(0)
    rng.shuffle(words)
    n_train_words = int(0.8 * len(words))
    np.random.seed(0)                                 # the reference's generator draws from the global numpy RNG
    out = {"train": ([], [], []), "test": ([], [], [])}
    for i, w in enumerate(words):
        split = "train" if i < n_train_words else "test"
        centres = kb.get_key_centers_for_word(w)
        if len(centres) < 2:
            continue
        proto = kb.get_word_prototype(w, T)
        for _ in range(PER_WORD):
            g = ref.keyboard.generate_minimum_jerk_trajectory(centres, T, include_midpoints=True,
                                                              offset_std=float(np.random.uniform(0.02, 0.05)))
            g = np.clip(np.asarray(g, np.float64), [-1, -1, 0], [1, 1, 1])
            if g.shape != (T, 3) or not np.isfinite(g).all():
                continue
            out[split][0].append(g)
            out[split][1].append(np.asarray(proto, np.float64))
            out[split][2].append(w)
    pack = {}
    for split, (g, p, w) in out.items():
        pack[f"{split}_gesture"] = np.stack(g).astype(np.float16)
        pack[f"{split}_prototype"] = np.stack(p).astype(np.float16)
        pack[f"{split}_word"] = np.array(w)
    path = os.path.join(ROOT, "tests", "golden", "realistic_gestures.npz")
    np.savez_compressed(path, **pack)
    print({k: v.shape for k, v in pack.items()}, f"{os.path.getsize(path) / 1e6:.2f} MB -> {path}")
    g = pack["train_gesture"].astype(np.float64)
    print("x range", g[..., 0].min(), g[..., 0].max(), "y", g[..., 1].min(), g[..., 1].max(), "t", g[..., 2].min(), g[..., 2].max(),
          "t monotone", bool((np.diff(g[..., 2], axis=1) >= -1e-3).all()))


if __name__ == "__main__":
    main()
