#!/usr/bin/env python
"""Golden VCFs for command-line option variants on the EXISTING golden profile files (run in the authoring container
only): the unmodified reference binary (oracle/_ref/popdel_ref, built by oracle/Makefile) is run with extra options and
its VCF is committed as tests/golden/<case>/opt_<name>.vcf. tests/golden/options.json lists the variants; the auxiliary
input files (coverage file, duplicate profile list) are written here and committed too.
Usage: python tests/golden/make_golden_options.py"""
import json
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.path.join(ROOT, "oracle", "_ref", "popdel_ref")

VARIANTS = [
    # per-read-group active-coverage caps (-A) on top of a changed default (-a)
    dict(case="highcov", name="A", list="profiles.txt", args=["-A", "opt_A_cov.txt", "-a", "150"],
         files={"opt_A_cov.txt": "rg0 0\nrg1 22\n"}),
    dict(case="highcov", name="A_n", list="profiles.txt", args=["-A", "opt_A_cov.txt", "-a", "150", "-n"], files={}),
    # a copy of a two-read-group profile under another name: conflicting read-group IDs resolved by -e; the list also
    # names one file twice (ignored with a warning, utils_popdel.h:879-912)
    dict(case="mixedrg", name="e", list="opt_e_profiles.txt", args=["-e"], copies={"opt_e_dup.profile": "sample00002.profile"},
         files={"opt_e_profiles.txt": "sample00000.profile\nsample00002.profile opt_e_dup.profile\nsample00000.profile\nsample00004.profile\n"}),
    # output of failed calls, tighter window cover, more EM iterations, unsmoothed histograms
    dict(case="mixedrg", name="F_c", list="profiles.txt", args=["-F", "-c", "0.9"], files={}),
    # regions of interest: cut inside a contig (window-wise output shows every window next to the region's ends, a region
    # that starts inside a deletion, several disjoint regions on two contigs (-r twice + -R file; OVERLAPPING regions make the reference write every record twice, which is not reproduced), a
    # region through a gap, a whole contig by name
    dict(case="basic", name="r_mid_n", list="profiles.txt", args=["-r", "chr21:100001-300000", "-n"], files={}),
    dict(case="basic", name="r_mid", list="profiles.txt", args=["-r", "chr21:100,001-300,000"], files={}),
    dict(case="basic", name="r_indel_n", list="profiles.txt", args=["-r", "chr21:283000-356000", "-n"], files={}),
    dict(case="basic", name="r_open", list="profiles.txt", args=["-r", "chr21:250001"], files={}),
    dict(case="twocontigs", name="r_multi_n", list="profiles.txt", args=["-r", "chrB:20001-60000", "-r", "chrA:50001-100000", "-R", "opt_rois.txt", "-n"],
         files={"opt_rois.txt": "chrA:110001-135000\nchrB:100001-140000\n"}),
    dict(case="twocontigs", name="r_contig", list="profiles.txt", args=["-r", "chrB"], files={}),
    dict(case="gap", name="r_gap_n", list="profiles.txt", args=["-r", "chr2:110001-790000", "-n"], files={}),
    # uncompressed profiles (-x), unzipped by the reference's own `popdel view -o`
    dict(case="highcov", name="x", list="opt_x_profiles.txt", args=["-x"],
         unzip={"opt_x_s0.profile": "sample00000.profile", "opt_x_s1.profile": "sample00001.profile", "opt_x_s2.profile": "sample00002.profile"},
         files={"opt_x_profiles.txt": "opt_x_s0.profile\nopt_x_s1.profile\nopt_x_s2.profile\n"}),
    dict(case="offset", name="u_t_p", list="profiles.txt", args=["-u", "-t", "6", "-p", "0.01", "-s", "0.5", "-f", "200"], files={}),
]


def main():
    out = []
    for v in VARIANTS:
        d = os.path.join(HERE, v["case"])
        for fn, text in v["files"].items():
            open(os.path.join(d, fn), "w").write(text)
        for dst, src in v.get("copies", {}).items():
            shutil.copyfile(os.path.join(d, src), os.path.join(d, dst))
        for dst, src in v.get("unzip", {}).items():
            subprocess.run([REF, "view", src, "-o", dst], check=True, cwd=d, stdout=subprocess.DEVNULL)
        vcf = f"opt_{v['name']}.vcf"
        subprocess.run([REF, "call", v["list"], "-o", vcf] + v["args"], check=True, cwd=d, stdout=subprocess.DEVNULL)
        n = sum(1 for line in open(os.path.join(d, vcf)) if not line.startswith("#"))
        print(f"{v['case']}/{vcf}: {n} records")
        out.append(dict(case=v["case"], name=v["name"], list=v["list"], args=v["args"], vcf=vcf, records=n))
    json.dump(out, open(os.path.join(HERE, "options.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
