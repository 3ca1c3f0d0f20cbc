"""Drop-in `AsyncTools` package (reference: /root/reference/AsyncTools/__init__.py:1): `import AsyncTools` then
`AsyncTools.AsyncPPO.X` / `AsyncTools.utils.f`, or `from AsyncTools.AsyncPPO import AsyncPPO, EnvVectorizer, VecMemory`."""
from . import utils  # noqa: F401
from . import AsyncPPO  # noqa: F401
