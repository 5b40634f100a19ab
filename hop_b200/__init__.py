"""Import alias: ``hop_b200`` is the importable name of the package whose sources live in
``hop-heterogeneous-topology-based-multimodal-entanglement-for-co-speech-gesture-generation_b200/``
(a directory name Python cannot import directly because of the hyphens)."""
import os as _os

_REAL = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      'hop-heterogeneous-topology-based-multimodal-entanglement-for-co-speech-gesture-generation_b200')
__path__ = [_REAL]
with open(_os.path.join(_REAL, '__init__.py')) as _f:
    exec(compile(_f.read(), _os.path.join(_REAL, '__init__.py'), 'exec'))
