"""Progress reporting hook kept for signature compatibility (progress/progressnotifier.py:28-79 in the reference):
``progress_notifier.iterator(iterable)`` wraps the work loop, either with tqdm or with GUI callbacks."""


class ProgressNotifier:
    def __init__(self):
        self._on_progress = None
        self._on_details = None
        self._use_tqdm = False

    @staticmethod
    def progress_notifier_tqdm():
        p = ProgressNotifier()
        p._use_tqdm = True
        return p

    def set_progress_report(self, callback):
        self._on_progress = callback

    def set_progress_detail(self, callback):
        self._on_details = callback

    def iterator(self, iterable):
        if self._use_tqdm:
            try:
                from tqdm import tqdm
                return tqdm(iterable)
            except ImportError:
                return iterable
        return self._callback_iter(iterable)

    def _callback_iter(self, iterable):
        total = len(iterable) if hasattr(iterable, '__len__') else None
        for i, item in enumerate(iterable):
            yield item
            if self._on_progress is not None and total:
                self._on_progress((i + 1) / total)
            if self._on_details is not None:
                self._on_details(i + 1, total)
