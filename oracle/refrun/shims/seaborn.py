"""Stand-in for the absent plotting package (lib/utils/metrics.py:5; plots only, no arithmetic)."""
