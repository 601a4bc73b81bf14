"""cuda_quantum_simulator_b200 — B200-native (sm_100a) state-vector engine behind the qsim API of
rylanmalarchick/cuda-quantum-simulator.  Python mirror of the C++ classes; all compute happens in
libqsim_b200.so (hand-written CUDA).  There is no CPU fallback."""
from ._lib import GATE_DTYPE, InvalidArgument, OutOfRange, QsimError, LIB_PATH
from .circuit import Circuit, GateType, create_bell_circuit, create_ghz_circuit, create_random_circuit
from .noise import (BatchedSimulator, DensityMatrixSimulator, NoiseChannel, NoiseModel, NoiseType, NoisySimulator)
from .simulator import CompiledCircuit, Simulator, jit_set_dual, jit_set_mode, jit_stats, jit_wait

__all__ = [
    "GATE_DTYPE", "InvalidArgument", "OutOfRange", "QsimError", "LIB_PATH", "Circuit", "GateType",
    "create_bell_circuit", "create_ghz_circuit", "create_random_circuit", "CompiledCircuit", "Simulator", "jit_set_dual", "jit_set_mode", "jit_stats", "jit_wait",
    "BatchedSimulator", "DensityMatrixSimulator", "NoiseChannel", "NoiseModel", "NoiseType", "NoisySimulator",
]
