"""Golden fixture for the GPU prototype / minimum-jerk kernels.

Run in the build container only:  python -m oracle.make_keyboard_golden
Calls the UNMODIFIED reference's QWERTYKeyboard (src/shared/keyboard.py:633-864) for a word list that covers the edge
cases (one key, repeated key, unknown characters only, long words): get_word_prototype(word, T) and
get_minimum_jerk_trajectory(word, T, include_midpoints, offset_std) for three settings, the noisy one after
np.random.seed(5) with the words in list order.  -> tests/golden/keyboard_golden.npz
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle.ref_loader import load_reference  # noqa: E402

EDGE = ["a", "aa", "qq", "??", "it's", "ok", "pm", "zq", "mississippi", "counterrevolutionaries", "qwertyuiopasdfghjklzxcvbnm"]


def main():
    ref = load_reference()
    kb = ref.keyboard.QWERTYKeyboard()
    words = []
    for line in open(os.path.join(ref.root, "dataset", "wordfreq.txt")):
        p = line.split()
        if len(p) == 2 and p[1].isalpha() and p[1].isascii():
            words.append(p[1].lower())
    rng = np.random.RandomState(1)
    words = EDGE + [words[i] for i in rng.choice(len(words), 89, replace=False)]
    out = {"words": np.array(words)}
    for T in (128, 37):
        out[f"proto_T{T}"] = np.stack([kb.get_word_prototype(w, T) for w in words])
    out["mj_mid_clean"] = np.stack([kb.get_minimum_jerk_trajectory(w, 128, True, 0.0) for w in words])
    out["mj_nomid_clean"] = np.stack([kb.get_minimum_jerk_trajectory(w, 128, False, 0.0) for w in words])
    np.random.seed(5)
    out["mj_mid_noise003"] = np.stack([kb.get_minimum_jerk_trajectory(w, 128, True, 0.03) for w in words])
    np.random.seed(6)
    out["mj_nomid_noise005_T64"] = np.stack([kb.get_minimum_jerk_trajectory(w, 64, False, 0.05) for w in words])
    out["key_centers"] = np.array([[kb.key_centers[c][0], kb.key_centers[c][1]] for c in "abcdefghijklmnopqrstuvwxyz"])
    path = os.path.join(ROOT, "tests", "golden", "keyboard_golden.npz")
    np.savez_compressed(path, **out)
    print({k: v.shape for k, v in out.items()}, f"{os.path.getsize(path) / 1e6:.2f} MB")


if __name__ == "__main__":
    main()
