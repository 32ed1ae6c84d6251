"""CPU oracle — TEST INFRASTRUCTURE ONLY (see oracle/README.md)."""
