# Top-level convenience Makefile: the reference's command line (README.md:19-37) keeps working.
#   make coo|csr|ell|sigma_c|cmrs|all     -> ./bin/<format>   (run from this directory, no arguments)
#   make databases                        -> databases/cant.mtx, databases/cant-sorted.mtx
#                                            (cant-SHAPED stand-ins: the real files are LFS stubs)
#   make lib | oracle | clean
PKG := opencl-spmv-algorithms_b200
TARGETS := coo csr ell sigma_c cmrs

.PHONY: all lib oracle databases clean host-sanitize gpu-sanitize $(TARGETS)
all: $(TARGETS)

lib:
	$(MAKE) -C $(PKG) lib tools

$(TARGETS): lib
	$(MAKE) -C $(PKG)/host BIN_DIR=$(CURDIR)/bin $@

oracle:
	$(MAKE) -C oracle

databases: lib
	@mkdir -p databases
	$(PKG)/tools/gen_mtx --order col --out databases/cant.mtx
	$(PKG)/tools/gen_mtx --order row --out databases/cant-sorted.mtx

# sanitizers: host code on the CPU, kernels on a GPU box
host-sanitize: lib
	$(MAKE) -C $(PKG)/host sanitize
gpu-sanitize: lib
	compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "formats_random or tuning_variant"
	compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "formats_random"

clean:
	$(MAKE) -C $(PKG) clean
	$(RM) -r bin
