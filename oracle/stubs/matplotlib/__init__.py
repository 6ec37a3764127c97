"""Import stub: /root/reference/libs/utils/metrics.py:5 imports pyplot for
commented-out plotting code only."""
