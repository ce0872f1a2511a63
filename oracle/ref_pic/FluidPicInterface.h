// Stand-in for SWMF share/ FluidPicInterface.h (un-vendored: Config.pl clones SWMFsoftware/share at HEAD), OURS, test
// infrastructure.  The reference's pic.h includes the header and a few fluid-coupler routines call its accessors; nothing on
// the ECSIM particle path (Lapenta2017, InitStencil, ProcessCell) does.  Every accessor returns 0: the coupler is off
// (fast-wave.input: CouplerMode=off).
#pragma once
#include <string>
#include <vector>
#include <array>
// types of the same library that pic.h names in declarations only
class Writer {};
typedef std::vector<std::array<double, 7>> VectorPointList;
template <class T>
class MDArray {};
class FluidPicInterface {
 public:
  int nG_D[3] = {0, 0, 0};
  double CellSize_BD[1][3] = {{0, 0, 0}}, BlockMin_BD[1][3] = {{0, 0, 0}}, BlockMax_BD[1][3] = {{0, 0, 0}};
  std::vector<Writer> writer_I;
#define AMPS_B200_STUB(name) \
  template <class... A>      \
  double name(A...) {        \
    return 0.0;              \
  }
  AMPS_B200_STUB(get_qom) AMPS_B200_STUB(writers_write) AMPS_B200_STUB(writers_init) AMPS_B200_STUB(set_doSaveBinary)
  AMPS_B200_STUB(set_State_BGV) AMPS_B200_STUB(readParam) AMPS_B200_STUB(pic_to_Mhd_Vec) AMPS_B200_STUB(getsRegion)
  AMPS_B200_STUB(getiRegion) AMPS_B200_STUB(get_nS) AMPS_B200_STUB(getSi2NoT) AMPS_B200_STUB(getQiSpecies) AMPS_B200_STUB(getPICUz)
  AMPS_B200_STUB(getPICUy) AMPS_B200_STUB(getPICUx) AMPS_B200_STUB(getPICUth) AMPS_B200_STUB(getPICRhoNum) AMPS_B200_STUB(getPICPzz)
  AMPS_B200_STUB(getPICPyz) AMPS_B200_STUB(getPICPyy) AMPS_B200_STUB(getPICPxz) AMPS_B200_STUB(getPICPxy) AMPS_B200_STUB(getPICPxx)
  AMPS_B200_STUB(getPICPpar) AMPS_B200_STUB(getPICP) AMPS_B200_STUB(getPICJz) AMPS_B200_STUB(getPICJy) AMPS_B200_STUB(getPICJx)
  AMPS_B200_STUB(getNo2SiT) AMPS_B200_STUB(getMiSpecies) AMPS_B200_STUB(getFluidStartZ) AMPS_B200_STUB(getFluidStartY)
  AMPS_B200_STUB(getFluidStartX) AMPS_B200_STUB(getBz) AMPS_B200_STUB(getBy) AMPS_B200_STUB(getBx) AMPS_B200_STUB(fixPARAM)
#undef AMPS_B200_STUB
};
