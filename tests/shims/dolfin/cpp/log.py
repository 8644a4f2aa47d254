class LogLevel:
    INFO, WARNING, ERROR, DEBUG, PROGRESS = 20, 30, 40, 10, 16


def set_log_level(level):
    return None
