from .progressnotifier import ProgressNotifier  # noqa: F401
