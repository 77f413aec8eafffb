"""ORACLE — test infrastructure only (see oracle/README.md). ctypes binding of libnpswf_oracle.so."""
from .oracle import *  # noqa: F401,F403
