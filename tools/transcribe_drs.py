#!/usr/bin/env python3
"""Transcribes pkg/cache/scheduler/fair_sharing_test.go:37 TestDominantResourceShare
(Go table literals) into tests/golden/drs_cases.json.  Run in the build container
only (needs /root/reference); the JSON is committed.  The reference is PARSED, not
executed (no Go toolchain here)."""
import json
import re
import sys

SRC = "/root/reference/pkg/cache/scheduler/fair_sharing_test.go"
RES = {"corev1.ResourceCPU": "cpu", "corev1.ResourceMemory": "memory"}


def res_name(tok):
    tok = tok.strip()
    return RES.get(tok, tok.strip('"'))


def parse_frq(block):
    out = []
    for m in re.finditer(r'\{Flavor:\s*"([^"]+)",\s*Resource:\s*([^}]+)\}:\s*([\d_]+)', block):
        out.append([m.group(1), res_name(m.group(2)), int(m.group(3).replace("_", ""))])
    return out


def parse_quotas(text):
    rgs = []
    for rg in re.split(r'ResourceGroup\(', text)[1:]:
        flavors = []
        parts = re.split(r'MakeFlavorQuotas\("([^"]+)"\)', rg)
        for i in range(1, len(parts), 2):
            fname, body = parts[i], parts[i + 1]
            ress = []
            for m in re.finditer(r'ResourceQuotaWrapper\("([^"]+)"\)((?:\.\w+\("[^"]*"\))*)\.Append\(\)', body):
                q = {"name": m.group(1), "nominal": "0", "borrowingLimit": None, "lendingLimit": None}
                for k, v in re.findall(r'\.(\w+)\("([^"]*)"\)', m.group(2)):
                    q[{"NominalQuota": "nominal", "BorrowingLimit": "borrowingLimit", "LendingLimit": "lendingLimit"}[k]] = v
                ress.append(q)
            flavors.append({"flavor": fname, "resources": ress})
        rgs.append(flavors)
    return rgs


def parse_cq(text):
    name = re.search(r'MakeClusterQueue\("([^"]+)"\)', text).group(1)
    cohort = re.search(r'\.\s*Cohort\("([^"]+)"\)', text)
    fw = re.search(r'FairWeight\(resource\.MustParse\("([^"]+)"\)\)', text)
    return {"name": name, "cohort": cohort.group(1) if cohort else None,
            "fairWeight": fw.group(1) if fw else None, "resourceGroups": parse_quotas(text)}


def parse_cohorts(text):
    out = []
    parts = re.split(r'MakeCohort\("([^"]+)"\)', text)
    for i in range(1, len(parts), 2):
        body = parts[i + 1]
        par = re.search(r'\.\s*Parent\("([^"]+)"\)', body)
        fw = re.search(r'FairWeight\(resource\.MustParse\("([^"]+)"\)\)', body)
        out.append({"name": parts[i], "parent": par.group(1) if par else None,
                    "fairWeight": fw.group(1) if fw else None, "resourceGroups": parse_quotas(body)})
    return out


def field(block, name, nxt):
    m = re.search(r'\n\t\t\t' + name + r':(.*?)(?=\n\t\t\t(?:' + "|".join(nxt) + r'):|\Z)', block, re.S)
    return m.group(1) if m else None


def main():
    src = open(SRC).read()
    lines = src.split("\n")
    start = next(i for i, l in enumerate(lines) if "func TestDominantResourceShare" in l)
    end = next(i for i, l in enumerate(lines) if i > start and l.startswith("\tfor name, tc := range cases"))
    body = "\n".join(lines[start:end])
    heads = [(m.start(), m.group(1)) for m in re.finditer(r'\n\t\t"([^"]+)": \{', body)]
    cases = {}
    for k, (pos, name) in enumerate(heads):
        block = body[pos:heads[k + 1][0] if k + 1 < len(heads) else len(body)]
        line = src[:src.index(block.strip("\n")[:40])].count("\n") + 1
        names = ["usage", "clusterQueue", "lendingClusterQueue", "cohorts", "flvResQ", "want"]
        c = {"source": f"pkg/cache/scheduler/fair_sharing_test.go:{line}"}
        c["usage"] = parse_frq(field(block, "usage", names) or "")
        c["flvResQ"] = parse_frq(field(block, "flvResQ", names) or "")
        c["clusterQueue"] = parse_cq(field(block, "clusterQueue", names))
        l = field(block, "lendingClusterQueue", names)
        c["lendingClusterQueue"] = parse_cq(l) if l else None
        co = field(block, "cohorts", names)
        c["cohorts"] = parse_cohorts(co) if co else []
        want = []
        for m in re.finditer(r'Name:\s*"([^"]+)",\s*NodeType:\s*(\w+),\s*DrName:\s*([^,]+),\s*DrValue:\s*([\w\.]+),[^\n]*\n\s*Borrowing:\s*(true|false)', field(block, "want", names)):
            v = m.group(4)
            val = (2**63 - 1) if "MaxInt" in v else int(v.replace("_", ""))
            want.append({"name": m.group(1), "cohort": m.group(2) == "nodeTypeCohort", "drName": res_name(m.group(3)),
                         "drValue": val, "borrowing": m.group(5) == "true"})
        c["want"] = want
        cases[name] = c
    json.dump(cases, open(sys.argv[1] if len(sys.argv) > 1 else "tests/golden/drs_cases.json", "w"), indent=1)
    print(len(cases), "cases;", sum(len(c["want"]) for c in cases.values()), "want rows")


if __name__ == "__main__":
    main()
