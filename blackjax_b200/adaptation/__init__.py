from .window_adaptation import build_schedule, window_adaptation  # noqa: F401
