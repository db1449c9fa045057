"""Shared-secret helpers for the launcher's RPC services (reference runner/common/util/secret.py)."""
import base64
import hashlib
import hmac
import os

SECRET_LENGTH = 32  # bytes
DIGEST_LENGTH = 32  # sha256
HOROVOD_SECRET_KEY = '_HOROVOD_SECRET_KEY'


def make_secret_key():
    return os.urandom(SECRET_LENGTH)


def encode_key(key):
    return base64.b64encode(key).decode('ascii')


def decode_key(text):
    return base64.b64decode(text.encode('ascii'))


def compute_digest(key, message):
    return hmac.new(key, message, hashlib.sha256).digest()


def check_digest(key, message, digest):
    return hmac.compare_digest(compute_digest(key, message), digest)
