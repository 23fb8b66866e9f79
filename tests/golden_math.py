"""Independent mathematics used to generate / re-derive golden vectors (no oracle, no product):
the augmented-generator matrix exponential of each linear compartment model."""
import numpy as np
from scipy.linalg import expm


def generator(kernel, p):
    """(A, input_state, state_names) of each rate-constant kernel (Appendix A.4 of SURVEY.md)."""
    if kernel == "one_compartment":
        ke, = p
        return np.array([[-ke]]), 0
    if kernel == "one_compartment_with_absorption":
        ka, ke = p
        return np.array([[-ka, 0], [ka, -ke]]), 1
    if kernel == "two_compartments":
        ke, kcp, kpc = p
        return np.array([[-(ke + kcp), kpc], [kcp, -kpc]]), 0
    if kernel == "two_compartments_with_absorption":
        ke, ka, kcp, kpc = p
        return np.array([[-ka, 0, 0], [ka, -(ke + kcp), kpc], [0, kcp, -kpc]]), 1
    if kernel == "three_compartments":
        k10, k12, k13, k21, k31 = p
        return np.array([[-(k10 + k12 + k13), k21, k31], [k12, -k21, 0], [k13, 0, -k31]]), 0
    if kernel == "three_compartments_with_absorption":
        ka, k10, k12, k13, k21, k31 = p
        return np.array([[-ka, 0, 0, 0], [ka, -(k10 + k12 + k13), k21, k31], [0, k12, -k21, 0], [0, k13, 0, -k31]]), 1
    raise KeyError(kernel)


def step(kernel, p, x, dt, rate):
    A, inp = generator(kernel, p)
    n = A.shape[0]
    M = np.zeros((n + 1, n + 1))
    M[:n, :n] = A
    M[inp, n] = rate
    return (expm(M * dt) @ np.append(x, 1.0))[:n]
