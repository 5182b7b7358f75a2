"""TEST INFRASTRUCTURE (oracle/).  The deterministic random-init weight generator lives in the package
(distilcodec_nabeel_b200/random_init.py) because bench.py's GPU arm needs synthetic weights too and the product arm
must not import anything from oracle/; this module re-exports it for the oracle-side tests and fixture scripts."""
from distilcodec_nabeel_b200.random_init import (DEFAULT_CONFIG, checksum, load_config, make_state_dict,  # noqa: F401
                                                 manifest, split_state_dict)
