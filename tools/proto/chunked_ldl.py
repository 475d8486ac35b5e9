"""Numpy prototype of the next factorisation step: a banded SPD system split into four
chunks = two nested twisted pairs around a middle separator, with the separator's coupling ("spike") carried
through the adjacent inner chunks as extra right-hand sides.  Verifies the algebra against a dense solve and
prints the length of the sequential pivot chain and the flops of every stage, i.e. what must run where:

    python tools/proto/chunked_ldl.py [n_blocks] [half_bandwidth_blocks]

Stages (T = rows above the middle separator S, B = rows below; T and B are each solved by the existing twisted
scheme, which is exact for a banded matrix):
  1. factorise T and B independently (4 chunk factorisations on 4 SMs + 2 inner separators)      chain: n/4 columns
  2. Y_T = T^-1 F_T^T, Y_B = B^-1 F_B^T  (180 right-hand sides each; only the chunk NEXT to S sees non-zeros
     before the inner separator) -> in CUDA: forward substitutions pipelined behind stage 1 on idle SMs
  3. S' = S - F_T Y_T - F_B Y_B ; r_S' = r_S - F_T T^-1 r_T - F_B B^-1 r_B                           SYRK-like, parallel
  4. x_S = S'^-1 r_S'                                                                                 30 columns
  5. x_T = T^-1 r_T - Y_T x_S ; x_B likewise                                                         parallel
"""
import sys
import numpy as np


def banded_spd(n, b, rng, bs=6):
    N = n * bs
    A = np.zeros((N, N))
    for i in range(n):
        for j in range(max(0, i - b), i + 1):
            blk = 0.3 * rng.uniform(-1, 1, (bs, bs))
            A[i * bs:(i + 1) * bs, j * bs:(j + 1) * bs] = blk
            A[j * bs:(j + 1) * bs, i * bs:(i + 1) * bs] = blk.T
        A[i * bs:(i + 1) * bs, i * bs:(i + 1) * bs] = (A[i * bs:(i + 1) * bs, i * bs:(i + 1) * bs] + A[i * bs:(i + 1) * bs, i * bs:(i + 1) * bs].T) / 2 + (4 * b + 10) * np.eye(bs)
    return A


def twisted_solve(A, rhs, b, bs=6):
    """The existing scheme: top half in natural order, bottom half reversed, joined at a separator of <= b blocks.
    Implemented with dense Schur complements (exact); rhs may have several columns.  Returns x and the pivot-chain
    length in block columns."""
    rhs2 = rhs.reshape(rhs.shape[0], -1)
    N = A.shape[0]
    n = N // bs
    if n <= 2 * b + 2:
        return np.linalg.solve(A, rhs2).reshape(rhs.shape), n
    m = n // 2
    top = slice(0, m * bs); sep = slice(m * bs, (m + b) * bs); bot = slice((m + b) * bs, N)
    Fst, Fsb = A[sep, top], A[sep, bot]
    k = Fst.shape[0]
    Yt = np.linalg.solve(A[top, top], np.column_stack([Fst.T, rhs2[top]]))
    Yb = np.linalg.solve(A[bot, bot], np.column_stack([Fsb.T, rhs2[bot]]))
    S = A[sep, sep] - Fst @ Yt[:, :k] - Fsb @ Yb[:, :k]
    rs = rhs2[sep] - Fst @ Yt[:, k:] - Fsb @ Yb[:, k:]
    xs = np.linalg.solve(S, rs)
    x = np.empty_like(rhs2)
    x[sep] = xs
    x[top] = Yt[:, k:] - Yt[:, :k] @ xs
    x[bot] = Yb[:, k:] - Yb[:, :k] @ xs
    return x.reshape(rhs.shape), max(m, n - m - b)


def four_chunk_solve(A, rhs, b, bs=6):
    N = A.shape[0]
    n = N // bs
    m = (n - b) // 2
    top = slice(0, m * bs); sep = slice(m * bs, (m + b) * bs); bot = slice((m + b) * bs, N)
    k = b * bs
    Fst, Fsb = A[sep, top], A[sep, bot]
    # stages 1+2: each half is solved by the twisted scheme for 1 + 180 right-hand sides (its own rhs and the spike)
    Yt, chain_t = twisted_solve(A[top, top], np.column_stack([rhs[top], Fst.T]), b, bs)
    Yb, chain_b = twisted_solve(A[bot, bot], np.column_stack([rhs[bot], Fsb.T]), b, bs)
    # stage 3-4
    S = A[sep, sep] - Fst @ Yt[:, 1:] - Fsb @ Yb[:, 1:]
    rs = rhs[sep] - Fst @ Yt[:, 0] - Fsb @ Yb[:, 0]
    xs = np.linalg.solve(S, rs)
    x = np.empty_like(rhs)
    x[sep] = xs
    x[top] = Yt[:, 0] - Yt[:, 1:] @ xs
    x[bot] = Yb[:, 0] - Yb[:, 1:] @ xs
    # in the twisted order of T the rows next to S are eliminated FIRST by T's reversed chunk, so the spike is
    # non-zero through that whole chunk and T's inner separator (and nowhere in T's natural-order chunk)
    spike_rows = (m - m // 2) + 0
    return x, max(chain_t, chain_b), k, spike_rows


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 400
    b = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    rng = np.random.default_rng(0)
    A = banded_spd(n, b, rng)
    rhs = rng.uniform(-1, 1, A.shape[0])
    x_ref = np.linalg.solve(A, rhs)
    x2, chain2 = twisted_solve(A, rhs, b)
    x4, chain4, k, spike_rows = four_chunk_solve(A, rhs, b)
    print(f"n = {n} block rows, half bandwidth {b} blocks")
    print(f"  twisted (2 chunks): |x - x_ref| = {np.abs(x2 - x_ref).max():.2e}, pivot chain {chain2} block columns")
    print(f"  four chunks       : |x - x_ref| = {np.abs(x4 - x_ref).max():.2e}, pivot chain {chain4} block columns "
          f"(+ {b} for the outer separator), spike = {k} right-hand sides over {spike_rows} block rows per side")
    upd = b * (b + 1) / 2 * 432            # FP64 flops of one pivot column's trailing update
    spike = b * 36 * 2 * k                 # forward substitution of the spike: b blocks x 6x6 x k right-hand sides per row
    print(f"  per column: trailing update {upd/1e3:.0f} kflop on the factorising SM, spike {spike/1e3:.0f} kflop "
          f"({spike/upd:.1f}x) -> pipelined on {max(1, round(spike/upd*3))} other SMs at a third of the load each")
    print(f"  Schur complement of the outer separator: {2 * 2 * k * k * spike_rows * 6 / 1e6:.0f} Mflop (parallel)")
    assert np.abs(x4 - x_ref).max() < 1e-9 and np.abs(x2 - x_ref).max() < 1e-9


if __name__ == "__main__":
    main()
