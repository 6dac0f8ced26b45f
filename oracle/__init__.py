"""ORACLE — test infrastructure only. See oracle/mpm_oracle.hpp for scope and parity status."""
