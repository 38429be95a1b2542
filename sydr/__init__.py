"""Import surface of the reference package, so that user code written against SyDR keeps its imports:

    from sydr.dsp.acquisition import PCPS, TwoCorrelationPeakComparison
    from sydr.dsp.tracking import EPL, DLL_NNEML, PLL_costa, BorreLoopFilter
    from sydr.channel.channel_l1ca_borre import ChannelL1CA
    from sydr.receiver.receiver_gps_l1ca import ReceiverGPSL1CA          (main.py:4-8 of the reference)

Every `sydr.X` that exists as `sydr_b200.X` IS that module (one module object under two names): there is no
second implementation here.  The three modules of the reference's main.py that lie outside the hot path
(terminal GUI, HTML report, logging set-up) are no-op stand-ins, enough for main.py's
statements to execute; nothing else of the reference's package is pretended to exist (ImportError as usual).
"""
import importlib
import importlib.abc
import importlib.util
import sys

_PREFIX, _TARGET = "sydr.", "sydr_b200."
_STANDINS = {"sydr.enlightengui", "sydr.io.visualisation", "sydr.logger"}
_OWN = _STANDINS | {"sydr.io"}           # real files of this package (sydr/io holds a stand-in next to the aliased sink)


class _AliasLoader(importlib.abc.Loader):
    def __init__(self, target):
        self.target = target

    def create_module(self, spec):
        return importlib.import_module(self.target)

    def exec_module(self, module):          # already executed under its own name
        pass


class _AliasFinder(importlib.abc.MetaPathFinder):
    def find_spec(self, fullname, path=None, target=None):
        if not fullname.startswith(_PREFIX) or fullname in _OWN:
            return None
        real = _TARGET + fullname[len(_PREFIX):]
        try:
            spec = importlib.util.find_spec(real)
        except (ImportError, ValueError):
            return None
        if spec is None:
            return None
        return importlib.util.spec_from_loader(fullname, _AliasLoader(real), is_package=spec.submodule_search_locations is not None)


if not any(isinstance(f, _AliasFinder) for f in sys.meta_path):
    sys.meta_path.insert(0, _AliasFinder())

from sydr_b200 import SydrError, __version__  # noqa: E402,F401
