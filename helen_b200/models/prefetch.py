"""In-process prefetch of batches for the predict loop.

With the native feed library the batch is filled by C++ threads that do not hold the interpreter lock, so a worker
PROCESS per reader - the reference's DataLoader(num_workers=N), dataloader_predict.py / predict_gpu.py:72-79 - only adds
a copy of every batch through shared memory (~2.7 GB/s, measured in tools/feed_rate.py).  `ThreadPrefetcher` runs the
dataset's __getitem__ on one background thread, `depth` batches ahead of the consumer.
"""
import queue
import threading


class ThreadPrefetcher(object):
    def __init__(self, dataset, depth=2, held=2):
        """`held`: batches the consumer keeps alive at a time (the predict loop: the one on the GPU and the one whose
        predictions are being written).  A dataset that recycles `ring` buffer sets needs depth + 1 + held of them: `depth`
        queued, one being filled, `held` in use."""
        self.dataset, self.depth = dataset, max(int(depth), 1)
        ring = int(getattr(dataset, "ring", 0) or 0)
        if 0 < ring < self.depth + 1 + held:
            raise ValueError(f"a ring of {ring} buffer sets is too small for a prefetch depth of {self.depth} with {held} batches in use")

    def __len__(self):
        return len(self.dataset)

    def __iter__(self):
        items = queue.Queue(maxsize=self.depth)
        stop = threading.Event()

        def produce():
            try:
                for index in range(len(self.dataset)):
                    if stop.is_set():
                        return
                    items.put((None, self.dataset[index]))
                items.put((None, None))
            except BaseException as error:             # handed to the consumer, which re-raises it
                items.put((error, None))

        worker = threading.Thread(target=produce, name="helen-feed-prefetch", daemon=True)
        worker.start()
        try:
            while True:
                error, item = items.get()
                if error is not None:
                    raise error
                if item is None:
                    return
                yield item
        finally:
            stop.set()
            while worker.is_alive():                   # unblock a producer waiting on a full queue
                try:
                    items.get_nowait()
                except queue.Empty:
                    worker.join(0.01)
