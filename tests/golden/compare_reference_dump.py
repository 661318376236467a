"""Compares the reference crate's own closest hits (rust/examples/golden_dump.rs, run with cargo where a Rust toolchain
exists) with the oracle's frozen answers in tests/golden/c2_hits_every97.npz.

    python tests/golden/compare_reference_dump.py c2_hits_every97.ref.bin

Record layout: u64 t_bits, u32 material, u32 n_candidates, u32 candidates[8] (48 bytes, little endian). Passing this check is
what turns "parity unpinned"  into a pin: t must agree bit for bit, and the oracle's leaf id must be one of the
primitives the reference itself reports at that t (the reference returns no leaf id, bvh.rs:121-124)."""
import os
import sys

import numpy as np

REC = np.dtype([("t_bits", "<u8"), ("material", "<u4"), ("n_cand", "<u4"), ("cand", "<u4", (8,))])


def compare(ref: np.ndarray, leaf: np.ndarray, t_bits: np.ndarray, n_triangles: int = 4968):
    """returns (n_t_mismatch, n_leaf_mismatch, n_material_mismatch)"""
    assert len(ref) == len(leaf) == len(t_bits), (len(ref), len(leaf))
    bad_t = int((ref["t_bits"] != t_bits).sum())
    miss = leaf == 0xFFFFFFFF
    in_cand = (ref["cand"] == leaf[:, None]).any(axis=1) & (leaf[:, None] != 0xFFFFFFFF).any(axis=1)
    overflow = ref["n_cand"] > 8  # more ties than the record holds: cannot decide, counted as agreeing
    bad_leaf = int((~miss & ~in_cand & ~overflow).sum()) + int((miss != (ref["material"] == 0xFFFFFFFF)).sum())
    mat = np.where(miss, 0xFFFFFFFF, np.where(leaf < n_triangles, 0, 1)).astype(np.uint32)  # bunny scene: triangles -> material 0, ground sphere -> 1
    bad_mat = int((mat != ref["material"]).sum())
    return bad_t, bad_leaf, bad_mat


def main():
    here = os.path.dirname(os.path.abspath(__file__))
    ref = np.fromfile(sys.argv[1], dtype=REC)
    z = np.load(os.path.join(here, "c2_hits_every97.npz"))
    bad_t, bad_leaf, bad_mat = compare(ref, z["leaf"], z["t_bits"])
    print(f"{len(ref)} records: t mismatches {bad_t}, leaf not among the reference's candidates {bad_leaf}, material mismatches {bad_mat}")
    sys.exit(0 if bad_t == 0 and bad_leaf == 0 and bad_mat == 0 else 1)


if __name__ == "__main__":
    main()
