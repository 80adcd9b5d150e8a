"""ANSI escape sequences behind the INFO / WARN / ERROR lines the command line prints; the attribute names are
the ones the reference's messages use (helen/modules/python/TextColor.py), so user scripts that strip or match
them keep working."""


def _sgr(code):
    return '\033[%dm' % code


class TextColor:
    END = _sgr(0)
    BOLD, UNDERLINE = _sgr(1), _sgr(4)
    DARKCYAN = _sgr(36)
    RED, GREEN, YELLOW, BLUE, PURPLE, CYAN = (_sgr(code) for code in range(91, 97))
