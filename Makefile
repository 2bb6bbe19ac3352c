.PHONY: metaseg meta_overlay build test clean

# same task name as the reference's Makefile:6-7
metaseg: build
	python src/metaseg.py

# same task name as the reference's Makefile
meta_overlay: build
	python src/meta_overlay.py

build:
	python -c "import __graft_entry__ as g; g.build()"

test:
	python -m pytest tests -x -q -m "not gpu"

clean:
	$(MAKE) -C ecseg_b200/csrc clean
	rm -rf oracle/_ref
