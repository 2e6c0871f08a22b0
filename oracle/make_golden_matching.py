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


def main():
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
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
