"""Corpus listing and split helpers with the reference's signatures (Code/data_processing.py:8,40).
No audio arithmetic happens here; they are kept importable so the reference's scripts can switch
packages without edits."""
import numpy as np


def load_esc(loc="../ESC-50-master/meta/esc50.csv", loc_audio="../ESC-50-master/audio/",
             list_categories=("dog", "chainsaw", "crackling_fire", "helicopter", "rain", "crying_baby",
                              "clock_tick", "sneezing", "rooster", "sea_waves")):
    """ESC-50 csv -> (audio paths ndarray, integer labels ndarray) restricted to ``list_categories``;
    label = position of the category in the list (Code/data_processing.py:8-38)."""
    import pandas as pd
    meta = pd.read_csv(loc)
    order = {c: i for i, c in enumerate(list_categories)}
    keep = meta[meta["category"].isin(order)]
    paths = np.array([loc_audio + f for f in keep["filename"]])
    labels = np.array([order[c] for c in keep["category"]])
    return paths, labels


def tt_split(list_audio_locs, l, f=0.8):
    """Per-class random train/test split, fraction f for training (Code/data_processing.py:40-65).
    Consumes the global numpy RNG once per class like the reference, so a fixed np.random.seed gives a
    reproducible split."""
    paths = np.asarray(list_audio_locs)
    labels = np.asarray(l)
    tr_p, tr_l, te_p, te_l = [], [], [], []
    for c in range(int(labels.max()) + 1):
        members = np.where(labels == c)[0]
        perm = np.random.permutation(len(members))
        n_train = int(f * len(members))
        for j, m in enumerate(members[perm]):
            (tr_p if j < n_train else te_p).append(paths[m])
            (tr_l if j < n_train else te_l).append(c)
    return tr_p, tr_l, te_p, te_l
