"""Drop-in `configs` package: `model_config` comes from this directory (same keys and values); every other module of the
reference's `configs/` (run_config.py with its paths, device, lr ...) keeps resolving to the reference checkout further
down sys.path -- the package path is extended over it."""
from pkgutil import extend_path

__path__ = extend_path(__path__, __name__)
