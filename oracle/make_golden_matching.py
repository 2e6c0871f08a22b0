"""Generates tests/golden/matching_*.pt by running the REFERENCE's own `match_features`
(/root/reference/ppeadepth/networks/replk_matching_adapter.py:261-340, unbound, on a stand-in self) on seeded synthetic
features.  Run in the build container (needs /root/reference):  python -m oracle.make_golden_matching"""
import os

import torch

from . import matching_oracle as M

CASES = {
    "matching_2x16x12x20_d8": dict(B=2, Fr=1, C=16, h=12, w=20, D=8, seed=0),
    "matching_2frames_missing_pose": dict(B=2, Fr=2, C=8, h=16, w=24, D=6, seed=1, zero_pose_item=1),
    "matching_far_bins": dict(B=1, Fr=1, C=4, h=10, w=14, D=12, seed=2, min_bin=0.05, max_bin=40.0),
}


# match_features_dyn (replk_matching_adapter.py:163-258): the reference method only runs at 48x128 with 96 bins (hard-coded, :166, :199),
# so the fixtures keep every KEPT_BINS-th layer of its (B,96,48,128) outputs plus float64 checksums of the whole volume.
DYN_CASES = {
    "matchdyn_cvmin_pool": (dict(B=1, C=4, seed=0), dict(cv_min=True, set_1=False, pool=True, pool_r=1, pool_th=0.7)),
    "matchdyn_avg_set1": (dict(B=1, C=4, seed=1), dict(cv_min=False, set_1=True, pool=False, pool_r=1, pool_th=0.7)),
    "matchdyn_cvmin_augmented": (dict(B=2, C=4, seed=2, augmented_item=1), dict(cv_min=True, set_1=False, pool=True, pool_r=1, pool_th=0.7)),
}
KEPT_BINS = 8


def make_dyn(out):
    import numpy as np
    for name, (kw, opts) in DYN_CASES.items():
        cur, look, poses, K, invK, bins, img, aug = M.synthetic_dyn_case(**kw)
        cost, missing = M.run_reference_match_features_dyn(cur, look, poses, K, invK, bins, img, aug_mask=aug, **opts)
        keep = list(range(0, cost.shape[1], KEPT_BINS))
        torch.save(dict(case=kw, opts=opts, cur=cur, look=look, poses=poses, K=K, invK=invK, bins=torch.as_tensor(bins),
                        images_u8=(img * 255).round().to(torch.uint8), aug=aug, kept_bins=keep, cost=cost[:, keep].clone(),
                        missing_bits=torch.from_numpy(np.packbits(missing.numpy().astype(np.uint8))),
                        cost_sum=float(cost.double().sum()), cost_sq=float((cost.double() ** 2).sum())),
                   os.path.join(out, name + ".pt"))
        print(name, tuple(cost.shape), "missing", float(missing.mean()))


def main():
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
    make_dyn(out)
    for name, kw in CASES.items():
        cur, look, poses, K, invK, bins = M.synthetic_case(**kw)
        for stm in (True, False):
            cost, missing = M.run_reference_match_features(cur, look, poses, K, invK, bins, set_missing_to_max=stm)
            torch.save(dict(case=kw, cur=cur, look=look, poses=poses, K=K, invK=invK, bins=torch.as_tensor(bins),
                            set_missing_to_max=stm, cost=cost, missing=missing),
                       os.path.join(out, "%s_%s.pt" % (name, "max" if stm else "raw")))
            print(name, stm, tuple(cost.shape), float(missing.mean()))


if __name__ == "__main__":
    main()
