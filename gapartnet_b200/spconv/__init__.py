"""Drop-in for the `spconv` package as GAPartNet uses it (`import spconv.pytorch as spconv`)."""
from . import pytorch  # noqa: F401
