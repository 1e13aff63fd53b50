// Stand-in for vtkXMLImageDataWriter: writes an uncompressed raw-appended
// .vti with the PointData array "ImageScalars" (Float64, N components), the
// same logical content VTK would emit for object2d.cpp:25-28.
// TEST INFRASTRUCTURE ONLY (see vtkSmartPointer.h).
#pragma once
#include <cstdint>
#include <cstdio>
#include <string>
#include <vtkImageData.h>

class vtkXMLImageDataWriter {
public:
    void SetFileName(const char* f) { _file = f; }
    void SetInputData(vtkImageData* img) { _img = img; }
    int Write() {
        if (!_img) return 0;
        FILE* fp = std::fopen(_file.c_str(), "wb");
        if (!fp) return 0;
        int* d = _img->GetDimensions();
        const auto& raw = _img->raw();
        const uint64_t nbytes = raw.size() * sizeof(double);
        std::fprintf(fp,
                     "<?xml version=\"1.0\"?>\n"
                     "<VTKFile type=\"ImageData\" version=\"1.0\" byte_order=\"LittleEndian\" "
                     "header_type=\"UInt64\">\n"
                     "  <ImageData WholeExtent=\"0 %d 0 %d 0 %d\" Origin=\"0 0 0\" Spacing=\"1 1 1\">\n"
                     "    <Piece Extent=\"0 %d 0 %d 0 %d\">\n"
                     "      <PointData Scalars=\"ImageScalars\">\n"
                     "        <DataArray type=\"Float64\" Name=\"ImageScalars\" NumberOfComponents=\"%d\" "
                     "format=\"appended\" offset=\"0\"/>\n"
                     "      </PointData>\n"
                     "      <CellData/>\n"
                     "    </Piece>\n"
                     "  </ImageData>\n"
                     "  <AppendedData encoding=\"raw\">\n_",
                     d[0] - 1, d[1] - 1, d[2] - 1, d[0] - 1, d[1] - 1, d[2] - 1, _img->components());
        std::fwrite(&nbytes, sizeof(nbytes), 1, fp);
        std::fwrite(raw.data(), 1, nbytes, fp);
        std::fprintf(fp, "\n  </AppendedData>\n</VTKFile>\n");
        std::fclose(fp);
        return 1;
    }

private:
    std::string _file;
    vtkImageData* _img = nullptr;
};
