"""Shadows the reference's ``losses`` package when dropin/ is first on sys.path."""
