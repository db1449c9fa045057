"""Shared secret of one job: authenticates the messages between the launcher's driver / task / notification services
(role parity: horovod/runner/common/util/secret.py).  HMAC-SHA256 over the message; the key travels to the workers in the
`_HOROVOD_SECRET_KEY` environment variable, base64-encoded."""
import base64
import hashlib
import hmac
import secrets

HOROVOD_SECRET_KEY = '_HOROVOD_SECRET_KEY'
_ALGO = hashlib.sha256
SECRET_LENGTH = 32                       # key bytes
DIGEST_LENGTH = _ALGO().digest_size      # 32


def make_secret_key() -> bytes:
    return secrets.token_bytes(SECRET_LENGTH)


def encode_key(key: bytes) -> str:
    return base64.b64encode(key).decode('ascii')


def decode_key(text: str) -> bytes:
    return base64.b64decode(text)


def compute_digest(key: bytes, message: bytes) -> bytes:
    mac = hmac.new(key, digestmod=_ALGO)
    mac.update(message)
    return mac.digest()


def check_digest(key: bytes, message: bytes, digest: bytes) -> bool:
    """Constant-time comparison of `digest` with the MAC of `message`."""
    return hmac.compare_digest(digest, compute_digest(key, message))
