"""B200-native retrieval hot path for A-NICE-RAG (dense scan + BM25 + weighted RRF + top-k).

The directory name carries a hyphen, so import it with
``importlib.import_module("a-nice-rag_b200")`` (see ``__graft_entry__.py``), or put
``a-nice-rag_b200/src`` on ``sys.path`` and use the reference's own module names:
``from search_engine import SearchEngine``, ``from database_manager import DatabaseManager``,
``from config import Config``.
"""
from . import native  # noqa: F401  (does not load the library yet)
from . import config, engine, registry  # noqa: F401
from .config import Config, InfoSource, SourceConfig  # noqa: F401
from .database_manager import DatabaseManager  # noqa: F401
from .search_engine import SearchEngine  # noqa: F401
from .batch_retrieval import retrieve_documents_batch  # noqa: F401
