"""import-only stub (blob_highlighter.py:15)."""
