"""Stand-in for sydr/logger.py (logging set-up from config/logging.ini, out of scope)."""
import logging


def configureLogger(name, filepath):
    return logging.getLogger(name)
