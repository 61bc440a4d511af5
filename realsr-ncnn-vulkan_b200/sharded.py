"""One image over several GPUs: tile-granular sharding.

The reference shards whole images over its GPUs (one RealSR object per GPU, proc threads pulling images from one queue,
/root/reference/src/main.cpp:778-828); a single large image (BASELINE configs[3]: 3840x2160 = 220 tiles at T=200) then
keeps one GPU busy.  Tiles share nothing but the read-only weights -- the 10-px halo is re-read from the source image,
src/realsr.cpp:180-181 -- so the tile grid is cut into contiguous row-major ranges of equal padded-pixel cost, each
engine (one per GPU, same model, same tilesize) runs `process_tiles` on its range from its own thread, and every GPU
copies its tiles' regions straight into the caller's output buffer: no collective, no assembly pass."""
import threading

from .realsr import RealSRError


def tile_costs(w, h, tilesize, prepadding):
    """padded network-input pixels of every tile, row-major over the reference tile grid (src/realsr.cpp:170-187)"""
    xt, yt = (w + tilesize - 1) // tilesize, (h + tilesize - 1) // tilesize
    costs = []
    for yi in range(yt):
        th = min((yi + 1) * tilesize, h) - yi * tilesize
        for xi in range(xt):
            tw = min((xi + 1) * tilesize, w) - xi * tilesize
            costs.append((tw + 2 * prepadding) * (th + 2 * prepadding))
    return costs


def split_tiles(w, h, tilesize, prepadding, nparts):
    """[(first, count)] * nparts: contiguous ranges covering every tile exactly once, cut where the running cost crosses
    k/nparts of the total (ranges may be empty when there are fewer tiles than parts)"""
    if nparts < 1:
        raise ValueError("nparts must be >= 1")
    costs = tile_costs(w, h, tilesize, prepadding)
    total = float(sum(costs))
    cuts, acc, k = [0], 0.0, 1
    for i, c in enumerate(costs):
        # tile i goes to part k-1; close the part once its share is reached (midpoint rule keeps the error <= half a tile)
        while k < nparts and acc + c / 2.0 >= total * k / nparts:
            cuts.append(i)
            k += 1
        acc += c
    while len(cuts) < nparts:
        cuts.append(len(costs))
    cuts.append(len(costs))
    return [(cuts[i], cuts[i + 1] - cuts[i]) for i in range(nparts)]


def process_sharded(engines, inimage, outimage):
    """engines: RealSR objects on different GPUs with the same model loaded; their tilesize/prepadding are taken from
    engines[0].  Returns 0 or the first non-zero code (message in that engine's last_error())."""
    if not engines:
        raise RealSRError("process_sharded needs at least one engine")
    e0 = engines[0]
    for e in engines[1:]:
        e.scale, e.tilesize, e.prepadding = e0.scale, e0.tilesize, e0.prepadding
    h, w, _ = inimage.shape
    parts = split_tiles(w, h, e0.tilesize, e0.prepadding, len(engines))
    rcs = [0] * len(engines)
    errs = [None] * len(engines)

    def work(i):
        first, count = parts[i]
        if not count:
            return
        try:
            rcs[i] = engines[i].process_tiles(inimage, outimage, first, count)
        except BaseException as ex:  # a worker thread must not swallow it: the shard's tiles would stay unwritten
            errs[i] = ex
            rcs[i] = -1

    threads = [threading.Thread(target=work, args=(i,)) for i in range(1, len(engines))]
    for t in threads:
        t.start()
    work(0)
    for t in threads:
        t.join()
    for ex in errs:
        if ex is not None:
            raise ex
    for rc in rcs:
        if rc:
            return rc
    return 0
