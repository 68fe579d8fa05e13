"""Dense exact diagonalisation of H = sum_ij J_ij sz_i sz_j - G sum_i sx_i - h sum_i sz_i
(the Hamiltonian implied by qmc_ising.rs:863-888 and the offset :97-99; `true` = sz=+1)."""
import numpy as np


def tfim_thermal(edges, nvars, transverse, longitudinal, beta):
    dim = 1 << nvars
    idx = np.arange(dim)
    sz = [1.0 - 2.0 * ((idx >> v) & 1) for v in range(nvars)]  # bit set = false?  choose bit=1 -> sz=-1
    sz = [-s for s in sz]  # bit=1 -> true -> sz=+1
    H = np.zeros((dim, dim))
    diag = np.zeros(dim)
    for (a, b), j in edges:
        diag += j * sz[a] * sz[b]
    for v in range(nvars):
        diag -= longitudinal * sz[v]
    H[idx, idx] = diag
    for v in range(nvars):
        H[idx, idx ^ (1 << v)] -= transverse
    w, U = np.linalg.eigh(H)
    w0 = w.min()
    bw = np.exp(-beta * (w - w0))
    Z = bw.sum()
    E = (w * bw).sum() / Z
    # <m^2>, <|m|> with m = mean sz (diagonal observable)
    m = sum(sz) / nvars
    probs = (U**2) @ (bw / Z)
    return {"E": E, "m": (m * probs).sum(), "m2": (m * m * probs).sum(), "absm": (np.abs(m) * probs).sum()}
