"""toolkit.models — same entry point as the reference (toolkit/models/__init__.py:29-70); only the SDUMC
model is provided (the reference's other 19 model files are absent from its own repository)."""
from sdumc_b200.model import WengnetMOSEIMultViewsTextMissing, get_models  # noqa: F401
