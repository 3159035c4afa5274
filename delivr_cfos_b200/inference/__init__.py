"""Mirror of the reference package ``inference`` (inference/inference.py, inference/sliding_window_inferer.py)."""
