# Convenience targets; the contract is __graft_entry__.py (build, smoke), tests/ and bench.py.
PY ?= python

.PHONY: build test test-gpu smoke bench bench-reference clean

build:            ## nvcc (sm_100a) + g++ -> cp-360-weakly-supervised-saliency_b200/lib/libcp360.so
	$(PY) -c "import __graft_entry__ as g; g.build()"

test: build       ## CPU suite: oracle vs fixtures / live reference, host builders, C-ABI, gloo sharding
	$(PY) -m pytest tests/ -x -q -m "not gpu"

test-gpu: build   ## GPU parity suite (needs a B200)
	$(PY) -m pytest tests/ -x -q -m gpu

smoke: build
	$(PY) -c "import __graft_entry__ as g; g.smoke()"

bench: build      ## one JSON line on stdout; GPUS=N shards frames over N GPUs
	$(PY) bench.py --gpus $(or $(GPUS),1)

bench-reference:  ## the reference's CPU path (port) on the host cores
	$(PY) bench.py --impl reference

clean:
	rm -rf cp-360-weakly-supervised-saliency_b200/lib cp-360-weakly-supervised-saliency_b200/build* .pytest_cache
