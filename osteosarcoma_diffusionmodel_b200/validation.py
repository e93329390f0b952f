"""GPU mirror of the three validators on the hot path (utils/validation.py:125-175, :177-223, :273-298)."""
