"""Environment registry (filled in as the env layer is built)."""
