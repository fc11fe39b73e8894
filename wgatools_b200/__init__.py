"""wgatools_b200 -- B200-native hot path of wgatools (maf2paf / stat / call): hand-written sm_100a CUDA
kernels behind a C ABI (include/wga_b200.h) with a host-side mirror of the reference's operator
interface (wgatools_b200.api).  No CPU fallback."""
from .api import (ChainReader, Comm, Context, Fasta, chain2maf, chain2paf, MAFReader, PAFReader, ResidentMaf, StatAcc, WGAError, call_var_maf, call_var_paf, default_context, generate_pesudo_maf,
                  filter_paf, filter_paf_align_pair, generate_maf_basedata, generate_paf_basedata, maf2chain, maf2paf, paf2chain, pafcov, parallel_validatepaf, pseudo_maf_files,
                  paf2maf, run_bytes, stat_maf, stat_paf)

__all__ = ["ChainReader", "Comm", "Context", "chain2maf", "chain2paf", "MAFReader", "PAFReader", "StatAcc", "WGAError", "call_var_maf", "call_var_paf", "default_context", "generate_pesudo_maf", "filter_paf", "filter_paf_align_pair", "generate_maf_basedata", "generate_paf_basedata", "maf2chain", "paf2chain", "pafcov", "parallel_validatepaf", "pseudo_maf_files",
           "maf2paf", "paf2maf", "Fasta", "ResidentMaf", "run_bytes", "stat_maf", "stat_paf"]
