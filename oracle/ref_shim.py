"""Import the UNMODIFIED reference (``/root/reference``) so its Python half can be executed as-is.

TEST INFRASTRUCTURE ONLY.  Used by ``oracle/make_golden.py`` (to generate ``tests/golden``) and by the
``not gpu`` tests that re-validate the restatement when ``/root/reference`` is mounted.  Nothing that
runs on the GPU box may call this (``/root/reference`` does not exist there).

Stubs needed: ``humanfriendly`` is not installed (only ``parse_size`` is used,
image.py:977,1013) and ``varKoder`` is not an installed distribution (core/config.py:15 asks
``importlib.metadata.version``).  ``dsk2ascii`` / ``reformat.sh`` do not exist here, so callers fake
``image.subprocess.run`` / ``image.run_parallel_reformats``.
"""
import importlib.metadata as _md
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("VARKODER_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "varKoder"))


def _parse_size(s):
    """decimal sizes as humanfriendly.parse_size does for K/M/G suffixes ("500K" -> 500000)."""
    s = str(s).strip()
    mult = {"K": 10**3, "M": 10**6, "G": 10**9, "T": 10**12}
    u = s.upper().rstrip("B")
    if u and u[-1] in mult:
        return int(float(u[:-1]) * mult[u[-1]])
    return int(float(u))


_loaded = {}


def load():
    """returns (image_module, utils_module, convert_module) of the reference."""
    if _loaded:
        return _loaded["image"], _loaded["utils"], _loaded["convert"]
    if not available():
        raise RuntimeError("reference not mounted at %s" % REFERENCE_ROOT)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    if "humanfriendly" not in sys.modules:
        hf = types.ModuleType("humanfriendly")
        hf.parse_size = _parse_size
        sys.modules["humanfriendly"] = hf
    real_version = _md.version

    def version(name):
        if name == "varKoder":
            return "1.4.0"          # pyproject.toml:7
        return real_version(name)

    _md.version = version
    try:
        import varKoder.commands.image as image
        import varKoder.core.utils as utils
        import varKoder.commands.convert as convert
    finally:
        _md.version = real_version
    _loaded.update(image=image, utils=utils, convert=convert)
    return image, utils, convert


class _FakeCompleted:
    def __init__(self, text):
        self.stdout = text.encode("UTF-8")
        self.stderr = b""


def reference_make_image(dsk2ascii_text, outfolder, kmer_mapping, sample="s", bp_k="00000010K", k=7,
                         mapping_code="varKode", labels=(), base_sd=0, subfolder_levels=0):
    """Run the reference ``make_image`` (image.py:808-936) with ``dsk2ascii`` replaced by ``dsk2ascii_text``.

    Returns (png_path, stats).  Only ``subprocess.run`` inside the reference module is faked.
    """
    from pathlib import Path
    image, _, _ = load()
    infile = Path(outfolder) / f"{sample}@{bp_k}+k{k}.fq.h5"
    real_run = image.subprocess.run
    image.subprocess.run = lambda *a, **kw: _FakeCompleted(dsk2ascii_text)
    try:
        stats = image.make_image(infile, Path(outfolder), kmer_mapping, overwrite=True, labels=list(labels),
                                 base_sd=base_sd, subfolder_levels=subfolder_levels, mapping_code=mapping_code)
    finally:
        image.subprocess.run = real_run
    return Path(outfolder) / f"{sample}@{bp_k}+{mapping_code}+k{k}.png", stats


def reference_ladder(nsites, min_bp, max_bp, is_query=False):
    """Run the reference ``split_fastq`` (image.py:629-725) on a synthetic gz file holding ``nsites`` bases
    with ``run_parallel_reformats`` stubbed; returns (sites_per_file, file names) or raises as it does."""
    import gzip
    import tempfile
    from pathlib import Path
    image, _, _ = load()
    captured = {}

    def fake_reformats(sites_per_file, outfs, infile, seed, verbose=False, max_workers=None):
        captured["sites"] = list(sites_per_file)
        captured["outfs"] = [Path(f).name for f in outfs]

    real = image.run_parallel_reformats
    image.run_parallel_reformats = fake_reformats
    try:
        with tempfile.TemporaryDirectory() as d:
            fq = Path(d) / "x.fq.gz"
            # one giant sequence line is enough: the reference only sums len(line)-1 over lines 1 mod 4
            with gzip.open(fq, "wb", compresslevel=1) as f:
                f.write(b"@r\n")
                chunk = b"A" * (1 << 20)
                left = nsites
                while left > 0:
                    f.write(chunk[: min(left, len(chunk))])
                    left -= min(left, len(chunk))
                f.write(b"\n+\n#\n")
            image.split_fastq(fq, "x", Path(d) / "out", min_bp=min_bp, max_bp=max_bp, is_query=is_query,
                              seed="1", overwrite=True)
    finally:
        image.run_parallel_reformats = real
    return captured["sites"], captured["outfs"]
