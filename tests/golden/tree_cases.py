"""Golden vectors transcribed BY HAND from the reference's table test
pkg/cache/scheduler/resource_test.go:31 TestAvailable (8 cases).  Each case keeps
the reference's builder calls (MakeClusterQueue(...).Cohort(...).ResourceGroup(
MakeFlavorQuotas(...).Resource(name, nominal, borrowingLimit, lendingLimit)))
and its `usage` / `wantAvailable` / `wantPotentiallyAvailable` int64 literals.
The test semantics (resource_test.go:355-420): with zero usage,
Available == PotentialAvailable == wantPotentiallyAvailable; after AddUsage(usage),
Available == wantAvailable and PotentialAvailable is unchanged.
"""
from kueue_b200.api import MakeClusterQueue as CQ, MakeCohort as Cohort, MakeFlavorQuotas as FQ

RC = ("red", "cpu")
BC = ("blue", "cpu")

TREE_CASES = {
    # resource_test.go:39
    "base cqs": dict(
        cqs=[CQ("cq1").ResourceGroup(FQ("red").Resource("cpu", "0")),
             CQ("cq2").ResourceGroup(FQ("red").Resource("cpu", "5"), FQ("blue").Resource("cpu", "10"))],
        cohorts=[],
        usage={"cq1": {RC: 1000}, "cq2": {RC: 2_500, BC: 1_000}},
        want_available={"cq1": {RC: 0}, "cq2": {RC: 2_500, BC: 9_000}},
        want_potential={"cq1": {RC: 0}, "cq2": {RC: 5_000, BC: 10_000}},
    ),
    # resource_test.go:67
    "cqs with cohort": dict(
        cqs=[CQ("cq1").Cohort("cohort").ResourceGroup(FQ("red").Resource("cpu", "10")),
             CQ("cq2").Cohort("cohort").ResourceGroup(FQ("red").Resource("cpu", "10", "", "9"))],
        cohorts=[Cohort("cohort").ResourceGroup(FQ("red").Resource("cpu", "10"))],
        usage={"cq1": {RC: 1000}, "cq2": {RC: 500}},
        want_available={"cq1": {RC: 28_000}, "cq2": {RC: 28_500}},
        want_potential={"cq1": {RC: 29_000}, "cq2": {RC: 30_000}},
    ),
    # resource_test.go:99
    "cq borrows from cohort": dict(
        cqs=[CQ("cq1").Cohort("cohort").ResourceGroup(FQ("red").Resource("cpu", "10"))],
        cohorts=[Cohort("cohort").ResourceGroup(FQ("red").Resource("cpu", "10"))],
        usage={"cq1": {RC: 11_000}},
        want_available={"cq1": {RC: 9_000}},
        want_potential={"cq1": {RC: 20_000}},
    ),
    # resource_test.go:117
    "cq oversubscription spills into cohort": dict(
        cqs=[CQ("cq1").Cohort("cohort").ResourceGroup(FQ("red").Resource("cpu", "10")),
             CQ("cq2").Cohort("cohort").ResourceGroup(FQ("red").Resource("cpu", "10", "0", "0"))],
        cohorts=[Cohort("cohort").ResourceGroup(FQ("red").Resource("cpu", "10"))],
        usage={"cq1": {RC: 31_000}},
        want_available={"cq1": {RC: 0}, "cq2": {RC: 0}},
        want_potential={"cq1": {RC: 20_000}, "cq2": {RC: 10_000}},
    ),
    # resource_test.go:154
    "lending and borrowing limits respected": dict(
        cqs=[CQ("cq1").Cohort("cohort").ResourceGroup(FQ("red").Resource("cpu", "10")),
             CQ("cq2").Cohort("cohort").ResourceGroup(FQ("red").Resource("cpu", "10", "2")),
             CQ("cq3").Cohort("cohort").ResourceGroup(FQ("red").Resource("cpu", "10", "", "5"))],
        cohorts=[Cohort("cohort").ResourceGroup(FQ("red").Resource("cpu", "10"))],
        usage={"cq1": {RC: 20_000}, "cq2": {RC: 10_000}, "cq3": {RC: 6_000}},
        want_available={"cq1": {RC: 4_000}, "cq2": {RC: 2_000}, "cq3": {RC: 4_000}},
        want_potential={"cq1": {RC: 35_000}, "cq2": {RC: 12_000}, "cq3": {RC: 40_000}},
    ),
    # resource_test.go:206
    "hierarchical cohort": dict(
        cqs=[CQ("cq1").Cohort("left").ResourceGroup(FQ("red").Resource("cpu", "10")),
             CQ("cq2").Cohort("right").ResourceGroup(FQ("red").Resource("cpu", "10", "", "0"))],
        cohorts=[Cohort("root"),
                 Cohort("left").Parent("root").ResourceGroup(FQ("red").Resource("cpu", "10")),
                 Cohort("right").Parent("root").ResourceGroup(FQ("red").Resource("cpu", "10"))],
        usage={"cq1": {RC: 10_000}, "cq2": {RC: 5_000}},
        want_available={"cq1": {RC: 20_000}, "cq2": {RC: 25_000}},
        want_potential={"cq1": {RC: 30_000}, "cq2": {RC: 40_000}},
    ),
    # resource_test.go:253
    "hierarchical cohort respects borrowing limit": dict(
        cqs=[CQ("left-cq1").Cohort("left").ResourceGroup(FQ("red").Resource("cpu", "10")),
             CQ("left-cq2").Cohort("left").ResourceGroup(FQ("red").Resource("cpu", "10", "5")),
             CQ("right-cq").Cohort("right").ResourceGroup(FQ("red").Resource("cpu", "10"))],
        cohorts=[Cohort("root"),
                 Cohort("left").Parent("root").ResourceGroup(FQ("red").Resource("cpu", "10", "5")),
                 Cohort("right").Parent("root").ResourceGroup(FQ("red").Resource("cpu", "10"))],
        usage={},
        want_available={"left-cq1": {RC: 35_000}, "left-cq2": {RC: 15_000}, "right-cq": {RC: 50_000}},
        want_potential={"left-cq1": {RC: 35_000}, "left-cq2": {RC: 15_000}, "right-cq": {RC: 50_000}},
    ),
    # resource_test.go:303
    "hierarchical cohort respects lending limit": dict(
        cqs=[CQ("left-cq1").Cohort("left").ResourceGroup(FQ("red").Resource("cpu", "10", "", "5")),
             CQ("left-cq2").Cohort("left").ResourceGroup(FQ("red").Resource("cpu", "10")),
             CQ("right-cq").Cohort("right").ResourceGroup(FQ("red").Resource("cpu", "0")),
             CQ("root-cq").Cohort("root").ResourceGroup(FQ("red").Resource("cpu", "0"))],
        cohorts=[Cohort("root"),
                 Cohort("left").Parent("root").ResourceGroup(FQ("red").Resource("cpu", "0", "", "5")),
                 Cohort("right").Parent("root").ResourceGroup(FQ("red").Resource("cpu", "0"))],
        usage={},
        want_available={"left-cq1": {RC: 20_000}, "left-cq2": {RC: 15_000}, "right-cq": {RC: 5_000}, "root-cq": {RC: 5_000}},
        want_potential={"left-cq1": {RC: 20_000}, "left-cq2": {RC: 15_000}, "right-cq": {RC: 5_000}, "root-cq": {RC: 5_000}},
    ),
}
