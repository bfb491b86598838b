"""gtsam.utils.plot placeholder (batch.py:27): plotting is out of scope."""


def __getattr__(name):
    def _unavailable(*args, **kwargs):
        raise NotImplementedError(f"gtsam.utils.plot.{name}: plotting is outside the batch LM path this package replaces")
    return _unavailable
