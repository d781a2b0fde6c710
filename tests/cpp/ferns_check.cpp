// Ferns on a synthetic sequence that leaves a view and comes back to it: the key-frame database fills, and findFrame at the
// revisit proposes the stored frame with a registration that matches the known relative pose.
// usage: ferns_check <raw.klg> <w> <h> <fx> <fy> <cx> <cy> <revisitTickOffset> [confidence]
#include <ElasticFusion.h>
#include <Ferns.h>
#include <Tools/RawLogReader.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>

int main(int argc, char** argv) {
  if (argc < 9) return 2;
  const int w = std::atoi(argv[2]), h = std::atoi(argv[3]);
  Resolution::getInstance(w, h);
  Intrinsics::getInstance((float)std::atof(argv[4]), (float)std::atof(argv[5]), (float)std::atof(argv[6]), (float)std::atof(argv[7]));
  const int tickOffset = std::atoi(argv[8]);
  // (a low confidence threshold lets the model prediction -- the smooth half of the fill-in views -- appear after a few frames)
  const float confidence = argc > 9 ? (float)std::atof(argv[9]) : 10.0f;
  ElasticFusion eFusion(2147483647 / 2, 35000, 5e-05, 1e-05, false, false, false, 115, confidence, 3, 10, false, 0.3095, true, false, "/tmp/ef_b200_ferns",
                        800000);
  Ferns ferns(500, 3000, 115, /*seed*/ 7);
  RawLogReader log(argv[1], false);
  int added = 0, frames = 0;
  while (log.hasMore()) {
    log.getNext();
    eFusion.processFrame(log.rgb, log.depth, log.timestamp, 1.0f);
    frames++;
    // ElasticFusion::processFerns (ElasticFusion.cpp:609-619): the fill-in views of the frame just fused
    GPUTexture fi(eFusion.context(), EF_BUF_FILL_IMAGE), fv(eFusion.context(), EF_BUF_FILL_VERTEX), fn(eFusion.context(), EF_BUF_FILL_NORMAL);
    added += ferns.addFrame(&fi, &fv, &fn, eFusion.get_T_wc(), eFusion.getTick() - 1, 0.3095f) ? 1 : 0;
  }
  std::printf("FRAMES %d ADDED %d STORED %zu\n", frames, added, ferns.frames.size());
  // query with the last frame's views, pretending `tickOffset` ticks have passed (findFrame only proposes frames older than 300)
  std::vector<Ferns::SurfaceConstraint> constraints;
  GPUTexture fi(eFusion.context(), EF_BUF_FILL_IMAGE), fv(eFusion.context(), EF_BUF_FILL_VERTEX), fn(eFusion.context(), EF_BUF_FILL_NORMAL);
  const ef::SE3d T = eFusion.get_T_wc();
  const ef::SE3d T_est = ferns.findFrame(constraints, T, &fv, &fn, &fi, eFusion.getTick() + tickOffset, false);
  const auto A = T.matrix();
  const auto B = T_est.matrix();
  double dmax = 0;
  for (int r = 0; r < 3; ++r) dmax = std::fmax(dmax, std::fabs((double)A(r, 3) - (double)B(r, 3)));
  for (size_t i = 0; i < ferns.frames.size(); ++i) {
    const auto F = ferns.frames[i]->T_wc.matrix();
    double d2 = 0;
    for (int r = 0; r < 3; ++r) d2 += ((double)A(r, 3) - (double)F(r, 3)) * ((double)A(r, 3) - (double)F(r, 3));
    std::fprintf(stderr, "stored %zu srcTime %d goodCodes %d dist_to_query %.4f\n", i, ferns.frames[i]->srcTime, ferns.frames[i]->goodCodes, std::sqrt(d2));
  }
  std::fprintf(stderr, "candidate %d dissimilarity %.4f\n", ferns.lastCandidate, (double)ferns.lastDissimilarity);
  std::printf("CLOSEST %d ICPERR %.6g ICPCOUNT %.0f PHOTO %.3f CONSTRAINTS %zu TDIFF %.5f\n", ferns.lastClosest, (double)ferns.lastICPError,
              (double)ferns.lastICPCount, (double)ferns.lastPhotoError, constraints.size(), dmax);
  return 0;
}
