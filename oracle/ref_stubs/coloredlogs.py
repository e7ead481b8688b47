"""Inert stand-in for coloredlogs (logconf.json:6 names coloredlogs.ColoredFormatter)."""
import logging


class ColoredFormatter(logging.Formatter):
    def __init__(self, fmt=None, datefmt=None, **k):
        super().__init__(fmt=fmt, datefmt=datefmt)
