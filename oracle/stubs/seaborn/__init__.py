"""Import stub: /root/reference/libs/utils/metrics.py:12 (unused at run time)."""
