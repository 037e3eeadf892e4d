"""Empty stand-in for `PyQt6`."""
from . import QtGui, QtCore, QtWidgets  # noqa: F401
