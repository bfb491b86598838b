"""gtsam.utils as imported by /root/reference/batch.py:27 (`from gtsam.utils import plot`).  Plotting is outside the path this
package replaces (batch.py:309-367); `plot` exists so the import succeeds and says so when used."""
from . import plot  # noqa: F401
