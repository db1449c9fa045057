"""Elastic rendezvous: the KV server answers `GET /rank_and_size/<host>:<local_rank>` by registering the worker as
READY (a barrier over the whole new world) and replying with its new `rank,size,local_rank,local_size,cross_rank,
cross_size,round`; `PUT /worker_addresses/<host>:<local_rank>` registers the worker's notification service.

Role parity: horovod/runner/elastic/rendezvous.py (create_rendezvous_handler) — installed here as scope hooks of
runner/http/http_server.py instead of a handler subclass.
"""
import logging

from horovod_b200.runner.common.util import network

# GET methods
GET_RANK_AND_SIZE = 'rank_and_size'
# PUT methods
PUT_WORKER_ADDRESSES = 'worker_addresses'


class ElasticRendezvousHandler(object):
    def __init__(self, driver):
        self._driver = driver

    def install(self, rendezvous_server):
        httpd = rendezvous_server.httpd
        httpd.get_hooks[GET_RANK_AND_SIZE] = self._get_rank_and_size
        httpd.put_hooks[PUT_WORKER_ADDRESSES] = self._put_worker_addresses

    def _get_rank_and_size(self, key, handler=None):
        host, local_rank = key.rsplit(':', 1)
        logging.info('_get_rank_and_size: {} {}'.format(host, local_rank))
        # a worker asking for its rank is (re-)initialising: it is READY for the next round
        rnd = self._driver.record_ready(host, int(local_rank))
        slot_info = self._driver.get_slot_info(host, int(local_rank))
        logging.info('rank and size: {} {}'.format(slot_info.rank, slot_info.size))
        return (slot_info.to_response_string() + ',' + str(rnd)).encode('ascii')

    def _put_worker_addresses(self, key, value):
        host, local_rank = key.rsplit(':', 1)
        addresses, secret_key = network.loads_base64(value.decode('ascii') if isinstance(value, bytes) else value)
        self._driver.register_worker_server(host, int(local_rank), addresses, secret_key)


def create_rendezvous_handler(driver):
    return ElasticRendezvousHandler(driver)
