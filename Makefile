# Convenience targets; the driver uses __graft_entry__.build() / pytest / bench.py directly.
PY ?= python
all: build
build:            ## libkbbq_b200.so (nvcc, sm_100a), the host mirror + CLI, the oracle (test infrastructure), test tools
	$(PY) -c "import __graft_entry__ as g; g.build()"
test:             ## CPU suite (no GPU needed)
	$(PY) -m pytest tests -q -m "not gpu"
test-gpu:         ## parity suite, on a B200
	$(PY) -m pytest tests -q -m gpu
bench:            ## one JSON line: config-2 shape on one GPU
	$(PY) bench.py
clean:
	$(MAKE) -C kbbq_b200/csrc clean
	$(MAKE) -C kbbq_b200/host clean
	$(MAKE) -C oracle clean
	rm -f tests/emu/libkbbq_emu.so
.PHONY: all build test test-gpu bench clean
