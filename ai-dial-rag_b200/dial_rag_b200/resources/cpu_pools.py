"""Executor pools of the boundary -- mirror of aidial_rag/resources/cpu_pools.py.

The threading contract is the reference's (cpu_pools.py:37-59): embedding calls arrive on
``indexing_embeddings`` / ``query_embeddings`` worker threads (1 worker each by default, so
one indexing and one query call may be inside the encoder at once -- the native encoder
serialises them with its own lock), exceptions raised in a worker propagate through the
awaited future (reference tests/test_cpu_pools.py:27-44).  Threads, not processes
(cpu_pools.py:43-44).
"""

from __future__ import annotations

import asyncio
import logging
import os
from concurrent.futures import ThreadPoolExecutor
from dataclasses import dataclass

CPU_COUNT = os.cpu_count() or 1
DEFAULT_CPU_POOL_WORKERS: int = max(1, CPU_COUNT - 2)

logger = logging.getLogger(__name__)


@dataclass
class CpuPoolsConfig:
    indexing_cpu_pool: int = DEFAULT_CPU_POOL_WORKERS
    indexing_embeddings_pool: int = 1
    query_embeddings_pool: int = 1


class CpuPools:
    def __init__(self, config: CpuPoolsConfig) -> None:
        self.indexing_cpu_pool = ThreadPoolExecutor(config.indexing_cpu_pool, thread_name_prefix="indexing_cpu")
        self.indexing_embeddings_pool = ThreadPoolExecutor(
            config.indexing_embeddings_pool, thread_name_prefix="indexing_embeddings")
        self.query_embeddings_pool = ThreadPoolExecutor(
            config.query_embeddings_pool, thread_name_prefix="query_embeddings")

    @staticmethod
    def _run_in_pool(pool, func, *args):
        return asyncio.get_running_loop().run_in_executor(pool, func, *args)

    def run_in_indexing_cpu_pool(self, func, *args):
        return self._run_in_pool(self.indexing_cpu_pool, func, *args)

    def run_in_indexing_embeddings_pool(self, func, *args):
        return self._run_in_pool(self.indexing_embeddings_pool, func, *args)

    def run_in_query_embeddings_pool(self, func, *args):
        return self._run_in_pool(self.query_embeddings_pool, func, *args)

    _instance = None

    @classmethod
    def instance(cls) -> "CpuPools":
        if cls._instance is None:
            logger.warning("CpuPools instance is not initialized. Initializing with default config.")
            cls.init_cpu_pools(CpuPoolsConfig())
        return cls._instance

    @classmethod
    def init_cpu_pools(cls, config: CpuPoolsConfig) -> "CpuPools":
        if cls._instance is not None:
            raise RuntimeError("CpuPools instance already initialized.")
        cls._instance = cls(config)
        return cls._instance


async def init_cpu_pools(config: CpuPoolsConfig):
    """Create and warm up the pools (first-call overhead), cpu_pools.py:99-105."""
    pools = CpuPools.init_cpu_pools(config)
    await pools.run_in_indexing_cpu_pool(sum, range(10))
    await pools.run_in_indexing_embeddings_pool(sum, range(10))
    await pools.run_in_query_embeddings_pool(sum, range(10))


def run_in_indexing_cpu_pool(func, *args):
    return CpuPools.instance().run_in_indexing_cpu_pool(func, *args)


def run_in_indexing_embeddings_pool(func, *args):
    return CpuPools.instance().run_in_indexing_embeddings_pool(func, *args)


def run_in_query_embeddings_pool(func, *args):
    return CpuPools.instance().run_in_query_embeddings_pool(func, *args)
