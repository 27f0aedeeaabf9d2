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


def _job_tag():
    """8 bytes every rank of THIS job derives identically (launcher run id, world size, port): a stray
    connection -- port scanner, health check, a rank of another job -- cannot present it."""
    import hashlib
    seed = "|".join([os.environ.get("PYLDA_JOB_NONCE", ""), os.environ.get("TORCHELASTIC_RUN_ID", ""),
                     os.environ.get("WORLD_SIZE", "1"), os.environ.get("MASTER_PORT", "29500")])
    return hashlib.sha256(seed.encode()).digest()[:8]


def _recv_exact(conn, nbytes):
    buf = b""
    while len(buf) < nbytes:
        chunk = conn.recv(nbytes - len(buf))
        if not chunk:
            break
        buf += chunk
    return buf


def exchange_unique_id(rank, size, make_id, timeout=300.0):
    """Rank 0 calls make_id() and serves the bytes to the size-1 other ranks; they connect (retrying
    until rank 0 listens), present `magic, rank, job tag` and receive the id.  Rank 0 listens on
    MASTER_ADDR only, ignores connections that do not present the handshake, counts distinct RANKS
    (a retry of the same rank is served again) and keeps serving until every rank has acknowledged.
    Returns the id on every rank."""
    if size <= 1:
        return make_id()
    host, port = _endpoint()
    tag = _job_tag()
    if rank == 0:
        uid = make_id()
        assert len(uid) == ID_BYTES
        srv = socket.socket(socket.AF_INET, socket.SOCK_STREAM)
        srv.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
        try:
            srv.bind((host, port))
        except OSError:
            srv.bind(("127.0.0.1", port))          # MASTER_ADDR that does not name a local interface
        srv.listen(size + 8)
        deadline = time.time() + timeout
        done = set()
        try:
            while len(done) < size - 1:
                srv.settimeout(max(0.1, deadline - time.time()))
                conn, _ = srv.accept()
                with conn:
                    conn.settimeout(5.0)
                    try:
                        hello = _recv_exact(conn, 16)
                        if len(hello) != 16 or hello[:4] != b"PLDA" or hello[8:] != tag:
                            continue                                  # not one of ours
                        peer = int.from_bytes(hello[4:8], "little")
                        if not 1 <= peer < size:
                            continue
                        conn.sendall(uid)
                        if _recv_exact(conn, 2) == b"OK":
                            done.add(peer)
                    except OSError:
                        continue
        finally:
            srv.close()
        return uid
    deadline = time.time() + timeout
    hello = b"PLDA" + int(rank).to_bytes(4, "little") + tag
    while True:
        try:
            with socket.create_connection((host, port), timeout=5.0) as c:
                c.sendall(hello)
                buf = _recv_exact(c, ID_BYTES)
                if len(buf) == ID_BYTES:
                    c.sendall(b"OK")
                    return buf
        except OSError:
            pass
        if time.time() > deadline:
            raise RuntimeError("pylda_b200: could not fetch the NCCL unique id from %s:%d" % (host, port))
        time.sleep(0.1)
