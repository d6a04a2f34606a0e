"""Request front-end for the batched engine — the engine side of /root/reference/serving.py:15-41.

The reference serves one request at a time through Flask (`Server.process_request`, serving.py:29-41; its call at
:33-37 passes `num_tokens` where `eos_token_id` belongs — stale against the current signature).  Flask is not part of this
image, so only what sits behind the route is built: the same `Server.process_request(request)` entry (argument order
fixed), `process_batch` for ragged batches, and `BatchingQueue`, which collects concurrent requests into batches for the
CUDA-graph engine (requests are independent: no state is shared between them beyond the models).  A Flask / FastAPI
route only has to call `queue.submit(request).result()`.
"""
from __future__ import annotations

import threading
import time
from concurrent.futures import Future
from typing import Callable, List, Optional, Sequence

import torch


class Server:
    """serving.py:15-41 with models passed in (no checkpoint download here) and an optional tokenizer: a request is
    {'prompt': str} (needs the tokenizer) or {'prompt_ids': [int, ...]}."""

    def __init__(self, approx_model, target_model, tokenizer=None, num_tokens: int = 40, top_k: int = 10,
                 top_p: float = 0.9, gamma: int = 4, temperature: float = 1.0, eos_token_id: Optional[int] = None,
                 pad_token_id: Optional[int] = None, random_seed: Optional[int] = None) -> None:
        self._small_model, self._large_model, self._tokenizer = approx_model, target_model, tokenizer
        dev = getattr(target_model, "device", None)
        if dev is None:
            try:
                dev = next(target_model.parameters()).device
            except (StopIteration, AttributeError):
                dev = "cuda"
        self._device = torch.device(dev)
        if self._device.type != "cuda":
            raise RuntimeError("serving.Server needs the models on a CUDA device: there is no CPU path")
        self.num_tokens, self.top_k, self.top_p = num_tokens, top_k, top_p          # serving.py:25-27
        self.gamma, self.temperature = gamma, temperature
        self.eos_token_id, self.pad_token_id, self.random_seed = eos_token_id, pad_token_id, random_seed
        self._served = 0

    def _ids(self, request: dict) -> torch.Tensor:
        if "prompt_ids" in request:
            return torch.as_tensor(request["prompt_ids"], dtype=torch.int64, device=self._device).reshape(-1)
        if self._tokenizer is None:
            raise ValueError("a {'prompt': str} request needs a tokenizer; send {'prompt_ids': [...]} instead")
        return self._tokenizer.encode(request["prompt"], return_tensors="pt").to(self._device).reshape(-1)   # serving.py:32

    def process_batch(self, requests: Sequence[dict]) -> List:
        """All requests of the list decode together (ragged prompts, one CUDA graph per iteration)."""
        from .sampling import speculative_sampling
        prompts = [self._ids(r) for r in requests]
        ids = list(range(self._served, self._served + len(prompts)))
        self._served += len(prompts)
        outs = speculative_sampling(prompts, self._small_model, self._large_model, self.eos_token_id, self.pad_token_id,
                                    self.num_tokens, gamma=self.gamma, temperature=self.temperature, top_k=self.top_k,
                                    top_p=self.top_p, random_seed=self.random_seed, request_ids=ids, pad_batch=True)
        res = []
        for r, o in zip(requests, outs):
            o = o.reshape(-1)
            if "prompt_ids" in r or self._tokenizer is None:
                res.append(o.tolist())
            else:
                res.append(self._tokenizer.decode(o, skip_special_tokens=True))                              # serving.py:38
        return res

    def process_request(self, request: dict):
        """serving.py:29-41 (one request)."""
        return self.process_batch([request])[0]


class BatchingQueue:
    """Collects concurrently submitted requests into batches: a worker thread takes up to `max_batch` requests, waiting
    at most `max_wait_s` for stragglers once the first one is there, and hands them to `handler(list) -> list`."""

    def __init__(self, handler: Callable[[List[dict]], List], max_batch: int = 64, max_wait_s: float = 0.005) -> None:
        self._handler, self._max_batch, self._max_wait = handler, int(max_batch), float(max_wait_s)
        self._items: List = []
        self._cv = threading.Condition()
        self._stop = False
        self.batches: List[int] = []                          # sizes of the batches served (statistics)
        self._thread = threading.Thread(target=self._run, daemon=True)
        self._thread.start()

    def submit(self, request: dict) -> Future:
        f: Future = Future()
        with self._cv:
            if self._stop:
                raise RuntimeError("queue is closed")
            self._items.append((request, f))
            self._cv.notify_all()
        return f

    def close(self) -> None:
        with self._cv:
            self._stop = True
            self._cv.notify_all()
        self._thread.join()

    def _run(self) -> None:
        while True:
            with self._cv:
                while not self._items and not self._stop:
                    self._cv.wait()
                if not self._items and self._stop:
                    return
                deadline = time.monotonic() + self._max_wait
                while len(self._items) < self._max_batch and not self._stop:
                    left = deadline - time.monotonic()
                    if left <= 0:
                        break
                    self._cv.wait(left)
                batch, self._items = self._items[:self._max_batch], self._items[self._max_batch:]
            try:
                results = self._handler([r for r, _ in batch])
                for (_, f), res in zip(batch, results):
                    f.set_result(res)
            except Exception as e:                            # noqa: BLE001 — every waiter must learn about the failure
                for _, f in batch:
                    f.set_exception(e)
            self.batches.append(len(batch))
