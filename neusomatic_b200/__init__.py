"""neusomatic_b200: B200-native candidate-scoring path of NeuSomatic."""
from . import _lib  # noqa: F401
from .engine import Engine, VARTYPE_CLASSES, anns_to_255  # noqa: F401
