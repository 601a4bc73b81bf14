"""Exact outcome distribution of BASELINE config C5 (12-qubit GHZ, depolarizing 0.005 + amplitude damping 0.001 on every
qubit after every gate: the BatchedSimulator schedule), from the oracle's exact Kraus evolution of the
4096 x 4096 density matrix (oracle/qsim_oracle.cpp: orc_dm_run_schedule).  Takes ~4 min of CPU, hence a committed
fixture: tests/golden/c5_exact_diag.npy (4096 float64).  Run: python tests/golden/make_c5_golden.py"""
import os
import sys
from ctypes import c_int64

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import helpers as H

N = 12
P_DEPOL, P_AD = 0.005, 0.001


def main():
    g = H.gates([("H", 0)] + [("CNOT", i, i + 1) for i in range(N - 1)])   # createGHZCircuit(12), src/Circuit.cpp:240-250
    ev = np.array([(t, qb, p) for t, p in ((0, P_DEPOL), (1, P_AD)) for qb in range(N)], H.CHANNEL_DTYPE)
    rho = np.zeros((1 << N, 1 << N), np.complex128)
    rho[0, 0] = 1
    rc = H.oracle().orc_dm_run_schedule(rho.ctypes.data_as(H.P), N, g.ctypes.data_as(H.P), c_int64(len(g)),
                                        ev.ctypes.data_as(H.P), c_int64(len(ev)))
    assert rc == 0
    diag = np.real(np.diag(rho)).copy()
    assert abs(diag.sum() - 1) < 1e-12
    np.save(os.path.join(HERE, "c5_exact_diag.npy"), diag)
    print("trace", diag.sum(), "p0", diag[0], "p_all_ones", diag[-1])


if __name__ == "__main__":
    main()
