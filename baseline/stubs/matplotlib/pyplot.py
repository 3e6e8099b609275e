"""stub: see matplotlib/__init__.py"""
