/* empty stand-in: popdel call/view never call htslib (oracle build shim) */
