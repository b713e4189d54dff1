"""audioblocks -- B200-native drop-in for the `audioblocks` package of
javierdrp/audio-effects-simulator (src/audioblocks/__init__.py:1-8 exports the same
13 names).  Put this directory's parent on sys.path in place of the reference's
`src/` and `src/backend.py` runs unchanged; the effect chain executes in
hand-written sm_100a CUDA kernels through the C ABI of libaesim.so."""
from .core import SmoothParam, EffectsChain, pick_devices, Effect, PlotDataTap, NativeEffect
from .delay import StereoDelayEffect
from .reverb import ReverbEffect
from .gate import NoiseGateEffect
from .spectral import SpectralFilter
from .octaver import OctaverEffect
from .filter import FilterEffect
from .distortion import DistortionEffect
from .convreverb import ConvolutionReverbEffect
from .engine import AudioEngine, SAMPLE_RATE
from . import sharding
from . import analysis
