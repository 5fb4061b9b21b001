"""TEST INFRASTRUCTURE — CPU oracle of the hot path.  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import this package; the product
(localization_b200/) never does."""
