"""neusomatic_b200: B200-native candidate-scoring path of NeuSomatic."""
from . import _lib  # noqa: F401


def __getattr__(name):
    # Engine and friends live in .engine, which imports torch: resolved on first use, so that the numpy-only modules
    # (vcf, predtable, fasta, synth, tsv) can be imported by the VCF stage's pool workers without loading torch
    if name in ("Engine", "VARTYPE_CLASSES", "anns_to_255"):
        from . import engine
        return getattr(engine, name)
    raise AttributeError("module %r has no attribute %r" % (__name__, name))
