"""Ordered, sequential batch map with progress lines -- mirror of aidial_rag/batched.py.

Semantics kept from the reference (batched.py:35-53): the input is cut into lists of
``batch_size``; batches are awaited ONE AT A TIME (no gather) so that heavy work of several
requests interleaves fairly on the single embeddings worker; results are flattened in
input order; a tqdm line is written to ``file`` at most every 10 s and at least every 30 s
as a keep-alive (batched.py:9-21).
"""

from __future__ import annotations

from itertools import chain, islice
from typing import Awaitable, Callable, Iterable, Iterator, List, TypeVar

from tqdm.std import tqdm as std_tqdm

T = TypeVar("T")
U = TypeVar("U")


class TqdmProgressBar(std_tqdm):
    def __init__(self, iterable=None, total=None, file=None):
        super().__init__(
            iterable=iterable, total=total, file=file,
            bar_format="{l_bar}{r_bar}\n",  # no bar glyphs; trailing newline for markdown
            mininterval=10, maxinterval=30, smoothing=0.5, position=0,
        )

    @staticmethod
    def status_printer(file):
        def print_status(s: str) -> None:
            file.write(s)  # no "\r": every update is its own line in the stage stream

        return print_status


def chunked(iterable: Iterable[T], n: int) -> Iterator[List[T]]:
    it = iter(iterable)
    while True:
        block = list(islice(it, n))
        if not block:
            return
        yield block


async def batched_map_with_progress(
    iterable: Iterable[T],
    coro_func: Callable[[List[T]], Awaitable[Iterable[U]]],
    batch_size: int,
    file,
) -> Iterable[U]:
    batches = list(chunked(iterable, batch_size))
    results = []
    for batch in TqdmProgressBar(iterable=batches, file=file):
        results.append(await coro_func(batch))  # strictly one batch in flight
    return chain.from_iterable(results)
