"""Same module path as the reference model file; the class is the sm_100a-backed drop-in."""
from sdumc_b200.model import WengnetMOSEIMultViewsTextMissing  # noqa: F401
