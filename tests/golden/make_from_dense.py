#!/usr/bin/env python
"""Builds tests/golden/decomposed_<case>_<mode>_from_dense.npz: a strictly feasible point of the clique-decomposed SDP
(built from the library's hand-off, tests/golden/handoff_<case>.npz) at the optimum of the DENSE LMI of the same
hand-off -- oracle/sdp_decomposed.certificate_from_dense: barrier solve of the dense problem, zero-fill LDL' of
-Z(gamma*) along the cliques of makeCliques (/root/reference/src/Methods/chordal_cliques.jl:13-59), split variables read
off the blocks -- together with the multipliers of the stored interior-point solution (decomposed_<case>_<mode>.npz) for
the dual side.  CPU only; W10-D20 takes about an hour on 8 cores.

    python tests/golden/make_from_dense.py W10-D20_beta2 single [precomputed.npz]
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
import sdp_decomposed as sd  # noqa: E402


def main():
    case, mode = sys.argv[1], sys.argv[2]
    h = np.load(os.path.join(HERE, f"handoff_{case}.npz"))
    cliques = sd.cliques_from_npz(h)
    if len(sys.argv) > 3:
        r = dict(np.load(sys.argv[3]))
    else:
        r = sd.certificate_from_dense(h, cliques, mode, verbose=True)
    ipm = np.load(os.path.join(HERE, f"decomposed_{case}_{mode}.npz"))
    U = max(float(r["U"]), float(ipm["U"]))      # the larger box contains both points; the multipliers belong to it
    assert float(ipm["U"]) >= float(r["U"]), "the stored multipliers were computed for a smaller box"
    np.savez_compressed(os.path.join(HERE, f"decomposed_{case}_{mode}_from_dense.npz"), mode=mode, U=U,
                        x=np.asarray(r["x"]), obj=float(r["obj"]), dense_obj=float(r["dense_obj"]),
                        X=np.asarray(ipm["X"]), xl=np.asarray(ipm["xl"]), xu=np.asarray(ipm["xu"]))
    print(f"{case} [{mode}]: objective {float(r['obj']):.10f} (dense barrier {float(r['dense_obj']):.10f}), "
          f"interior-point solution {float(ipm['obj']):.10f}, its dual bound {float(ipm['dual_obj']):.10f}")


if __name__ == "__main__":
    main()
