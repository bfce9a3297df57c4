"""Retriever weights and data locations the drop-in classes are configured with.

The orchestrators written against the reference read three things from its ``src/config.py``
(:7-59) -- ``InfoSource``, ``Config.DEFAULT_MODEL_WEIGHTS`` / ``Config.SOURCE_CONFIGS`` and
``Config.get_source_config`` -- so the same names, values and error behaviour are provided here
(``tests/test_host.py`` compares every public value with the reference's module where it is
mounted).  The tables below are the single place they are stated.
"""
import enum
from typing import Dict, Optional

# retriever name -> weight in the weighted RRF; a weight <= 0 switches the retriever off at the
# orchestrator (query_rag_retrieval.py:202-304).  Dense model first, BM25 last: 5 : 1.
_RETRIEVER_WEIGHTS = (
    ("voyage-3-large", 5.0),
    ("text-embedding-3-large", 0.0),
    ("voyage-3.5", 0.0),
    ("Qwen3", 0.0),
    ("BM25", 1.0),
)

_DATA_DIR = "databases"
# attribute of SourceConfig -> file under _DATA_DIR, per information source
_NICE_FILES = {
    "voyage_db_path": "voyage_3_large_nice_guidelines_2048.db",
    "voyage_3_5_db_path": "voyage_3.5_nice_guidelines_2048.db",
    "openai_db_path": "text_embedding_3_large_nice_guidelines.db",
    "qwen_db_path": "Qwen3-Embedding-0.6B_nice_guidelines.db",
    "bm25_path": "bm25_index_nice_guidelines.pkl",
}


class InfoSource(enum.Enum):
    NICE = "nice"


class SourceConfig:
    """Where one information source keeps its embedding databases and its BM25 index, and how the
    prompts talk about it.  ``voyage_db_path`` defaults to ``db_path`` (the primary database)."""

    __slots__ = ("db_path", "bm25_path", "context_description", "not_found_message",
                 "voyage_db_path", "voyage_3_5_db_path", "openai_db_path", "qwen_db_path")

    def __init__(self, db_path: str, bm25_path: str, context_description: str,
                 not_found_message: str, voyage_db_path: Optional[str] = None,
                 voyage_3_5_db_path: Optional[str] = None, openai_db_path: Optional[str] = None,
                 qwen_db_path: Optional[str] = None):
        self.db_path = db_path
        self.bm25_path = bm25_path
        self.context_description = context_description
        self.not_found_message = not_found_message
        self.voyage_db_path = db_path if voyage_db_path is None else voyage_db_path
        self.voyage_3_5_db_path = voyage_3_5_db_path
        self.openai_db_path = openai_db_path
        self.qwen_db_path = qwen_db_path

    def _fields(self):
        return tuple(getattr(self, name) for name in self.__slots__)

    def __eq__(self, other):
        return isinstance(other, SourceConfig) and self._fields() == other._fields()

    def __repr__(self):
        return "SourceConfig(" + ", ".join(f"{n}={getattr(self, n)!r}" for n in self.__slots__) + ")"


def _nice_source() -> SourceConfig:
    paths = {attr: f"{_DATA_DIR}/{name}" for attr, name in _NICE_FILES.items()}
    return SourceConfig(db_path=paths["voyage_db_path"], context_description="NICE guidelines",
                        not_found_message="no relevant NICE guidelines were found", **paths)


class Config:
    DEFAULT_MODEL_WEIGHTS: Dict[str, float] = dict(_RETRIEVER_WEIGHTS)
    SOURCE_CONFIGS: Dict[InfoSource, SourceConfig] = {InfoSource.NICE: _nice_source()}

    @classmethod
    def get_source_config(cls, source: str) -> SourceConfig:
        known = {member.value: member for member in InfoSource}
        member = known.get(source.lower())
        if member is None:
            raise ValueError(f"Unknown source: {source}. Valid sources: {list(known)}")
        return cls.SOURCE_CONFIGS[member]
