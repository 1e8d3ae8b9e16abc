"""Stand-in for the absent plotting package (lib/utils/metrics.py:4 imports it for draw_add_curve only; no arithmetic)."""
