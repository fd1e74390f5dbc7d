"""Reference-facing module tree of univl_b200 (same module names as the reference `modules/` package)."""
