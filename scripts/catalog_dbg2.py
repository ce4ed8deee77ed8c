import sys, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from recguru_b200 import ops
from oracle import ref_path as O
def grid(shape, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randint(-32, 33, shape, generator=g).float() / 16.0
U, N, k, d = 100, 130, 10, 128
users, items = grid((U, d), 10 + U), grid((N, d), 20 + N)
g = torch.Generator().manual_seed(U + N)
target = torch.randint(1, N + 1, (U,), generator=g)
scores = O.catalog_scores(users, items)
ub, ib = users.to(torch.bfloat16).cuda(), items.to(torch.bfloat16).cuda()
ts = ops.raw_catalog_target_scores(ub, ib, target.cuda(), 1)
val, idx, cnt = ops.raw_catalog_topk_rank(ub, ib, target.cuda(), ts, k, 1)
rv, ri = O.catalog_topk(scores, k)
bad = (idx.cpu() != ri + 1).any(1).nonzero().view(-1)
print("bad rows", bad.tolist()[:20], "of", U)
for r in bad[:4].tolist():
    print("row", r); print(" got ids ", idx[r].tolist()); print(" ref ids ", (ri[r] + 1).tolist())
    print(" got val ", val[r].tolist()); print(" ref val ", rv[r].tolist())
