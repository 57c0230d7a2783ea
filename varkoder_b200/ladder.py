"""Host-side ladder and file-name logic of ``split_fastq`` (reference: varKoder/commands/image.py:669-709).

Integer-only restatement; the device computes the same ladder in ``plan_kernel`` (csrc/vk_parse.cuh) so that
the fused path needs no host round trip.  The reference uses ``int(math.log10(x))`` and a float division,
which agree with the integer form for every x < 10**15.
"""

SAMPLE_BP_SEP = "@"      # core/config.py:21
BP_KMER_SEP = "+"        # core/config.py:20
LABELS_SEP = ";"         # core/config.py:19
QUAL_THRESH = 0.01       # core/config.py:24
MAX_LEVELS = 64


class LessThanMinimumData(Exception):
    """split_fastq raises a bare Exception with this message (image.py:680); kept as a subclass."""

    def __init__(self):
        super().__init__("Input file has less than minimum data.")


def ladder(nsites, min_bp=50000, max_bp=None, is_query=False):
    """``sites_per_file`` of split_fastq for a file holding ``nsites`` bases (image.py:669-695)."""
    nsites = int(nsites)
    min_bp = int(min_bp)
    if max_bp is None:
        sites = [nsites]
    elif is_query or nsites > min_bp:
        sites = [min(nsites, int(max_bp))]
    else:
        raise LessThanMinimumData()
    if not is_query:
        while sites[-1] > min_bp and len(sites) < MAX_LEVELS:
            oneless = sites[-1] - 1
            if oneless == 0:
                break
            p10 = 10 ** (len(str(oneless)) - 1)
            first_digit = oneless // p10
            mult = 5 if first_digit >= 5 else (2 if first_digit >= 2 else 1)
            sites.append(mult * p10)
        if sites[-1] < min_bp:
            del sites[-1]
    return sites


def level_tag(bp):
    """'%08dK' of the level's TARGET bases (image.py:704-705)."""
    return str(int(bp / 1000)).rjust(8, "0") + "K"


def image_name(sample, bp, mapping_code, k):
    """``sample@NNNNNNNNK+mapping+k7.png`` (image.py:699-709, 752-758, 843-849)."""
    return f"{sample}{SAMPLE_BP_SEP}{level_tag(bp)}{BP_KMER_SEP}{mapping_code}{BP_KMER_SEP}k{k}.png"


def parse_seed(seed):
    """split_fastq gets ``seed=str(row_index)+str(rng.integers(2**32))`` and reformat.sh receives
    ``int(seed)+i`` (image.py:585, 1017); the GPU path reduces it mod 2^64."""
    if seed is None:
        return 0
    return int(seed) & ((1 << 64) - 1)
