// Forwarding header: keeps the reference's include name for drop-in callers.
#pragma once
#include "qsim/optimized_gates.cuh"
