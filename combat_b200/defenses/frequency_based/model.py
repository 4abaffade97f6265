"""defenses/frequency_based/model.py of the reference (inference path)."""
from ...modules import FrequencyModel  # noqa: F401
