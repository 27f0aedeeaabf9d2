"""One process per GPU: rank discovery and the one host-side exchange the multi-GPU path needs --
shipping rank 0's 128-byte NCCL unique id to the other ranks.  Standard library only (the product
path stays free of torch); the data path is the NCCL all-reduce inside libpylda_b200.so.

Environment (what `torchrun` / `python -m torch.distributed.run` exports): RANK, WORLD_SIZE,
LOCAL_RANK, MASTER_ADDR, MASTER_PORT.  The id is served on MASTER_PORT + PYLDA_ID_PORT_OFFSET
(default 173) so that it does not collide with a rendezvous store already listening on MASTER_PORT.
"""
import os
import socket
import time

ID_BYTES = 128


def world():
    """(rank, world_size, local_rank) from the environment; (0, 1, PYLDA_DEVICE or 0) when not launched
    as a multi-process job."""
    rank = int(os.environ.get("RANK", "0"))
    size = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", os.environ.get("PYLDA_DEVICE", "0")))
    return rank, size, local


def _endpoint():
    host = os.environ.get("MASTER_ADDR", "127.0.0.1")
    port = int(os.environ.get("MASTER_PORT", "29500")) + int(os.environ.get("PYLDA_ID_PORT_OFFSET", "173"))
    return host, port


def exchange_unique_id(rank, size, make_id, timeout=300.0):
    """Rank 0 calls make_id() and serves the bytes to the size-1 other ranks; they connect (retrying
    until rank 0 listens) and receive them.  Returns the id on every rank."""
    if size <= 1:
        return make_id()
    host, port = _endpoint()
    if rank == 0:
        uid = make_id()
        assert len(uid) == ID_BYTES
        srv = socket.socket(socket.AF_INET, socket.SOCK_STREAM)
        srv.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
        srv.bind(("", port))
        srv.listen(size)
        srv.settimeout(timeout)
        served = 0
        try:
            while served < size - 1:
                conn, _ = srv.accept()
                with conn:
                    conn.sendall(uid)
                served += 1
        finally:
            srv.close()
        return uid
    deadline = time.time() + timeout
    while True:
        try:
            with socket.create_connection((host, port), timeout=5.0) as c:
                buf = b""
                while len(buf) < ID_BYTES:
                    chunk = c.recv(ID_BYTES - len(buf))
                    if not chunk:
                        break
                    buf += chunk
            if len(buf) == ID_BYTES:
                return buf
        except OSError:
            pass
        if time.time() > deadline:
            raise RuntimeError("pylda_b200: could not fetch the NCCL unique id from %s:%d" % (host, port))
        time.sleep(0.1)
