"""Mirror of nerfactor/util/logging.py:21-87: coloured console messages with a `[loggee]`
prefix and printf-style arguments (`logger.info("Epoch %d", i)`)."""
import logging as _pylog

_ANSI = {'red': 31, 'green': 32, 'pink': 35, 'cyan': 36}


class Logger:
    def __init__(self, prefix="", suffix="", loggee=None, debug_mode=False, use_absl=False):
        self.prefix = prefix + ("[%s] " % loggee if loggee is not None else "")
        self.suffix = suffix
        self.debug_mode = debug_mode
        self.use_absl = use_absl          # here: route to the std `logging` module instead

    def _emit(self, level, color, args, always=True):
        msg = self.prefix + (args[0] % tuple(args[1:])) + self.suffix
        if self.use_absl:
            _pylog.log(level, msg)
        elif always:
            print("\x1b[%dm%s\x1b[0m" % (_ANSI[color], msg))

    def info(self, *args, color='cyan'):
        self._emit(_pylog.INFO, color, args)

    def warn(self, *args, color='pink'):
        self._emit(_pylog.WARNING, color, args)

    warning = warn

    def error(self, *args, color='red'):
        self._emit(_pylog.ERROR, color, args)

    def debug(self, *args, color='green'):
        self._emit(_pylog.DEBUG, color, args, always=self.debug_mode)
