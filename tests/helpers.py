"""shared test helpers: hand-made FASTQ edge cases and oracle glue (oracle/ is only ever used as the checker)."""
import numpy as np

from oracle import dsk, image as oimg


def fastq(reads, quals=None, final_newline=True, headers=None):
    out = []
    for i, r in enumerate(reads):
        h = headers[i] if headers else f"@r{i}"
        q = quals[i] if quals else "I" * len(r)
        out.append(f"{h}\n{r}\n+\n{q}\n")
    s = "".join(out).encode("ascii")
    return s if final_newline else s[:-1]


def rand_reads(rng, n, lmin, lmax, p_n=0.01, alphabet="ACGT"):
    reads = []
    for _ in range(n):
        L = int(rng.integers(lmin, lmax + 1))
        a = rng.choice(list(alphabet), L)
        if p_n:
            a[rng.random(L) < p_n] = "N"
        reads.append("".join(a))
    return reads


def oracle_levels(buf, k, seed, levels, nsites, read_index_base=0):
    """canon_full per level with the project's selection rule, computed by the CPU oracle."""
    p = dsk.parse_fastq(buf)
    out = []
    for bp in levels:
        sel = dsk.select_reads(p["n_reads"], seed, bp, nsites, read_index_base)
        fwd = dsk.count_forward(buf, p["starts"], p["lens"], k, sel)
        out.append(dsk.fold_canonical(fwd, k))
    return np.stack(out) if out else np.zeros((0, 4 ** k), dtype=np.uint64)


def oracle_images(canon_levels, lut):
    return np.stack([oimg.image_exact(c, lut) for c in canon_levels])
